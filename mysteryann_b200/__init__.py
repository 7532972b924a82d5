"""roargraph-b200: the RoarGraph (matchyc/mysteryann) search / build-kNN hot path as sm_100a CUDA kernels behind a
C ABI (include/roargraph_b200.h).  The package holds the CUDA sources (csrc/), the drop-in host C++ layer (host/)
and small Python harness helpers (formats, synthetic data, ctypes binding)."""
__all__ = ["io", "synth", "capi", "build"]
