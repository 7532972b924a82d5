"""TEST INFRASTRUCTURE ONLY - ctypes bindings for the two checker libraries.

* `Oracle`  -> oracle/liboracle.so               (plain-C restatement, always available; `make -C oracle oracle`)
* `Ref`     -> oracle/_ref/libroargraph_ref.so   (the unmodified reference TUs; built where /root/reference
               exists, travels to the GPU box as a prebuilt .so; needs an AVX-512F/DQ host CPU)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libroargraph_ref.so")

_u8p = C.c_void_p


def _ptr(a):
    return None if a is None else a.ctypes.data


def build(ref: bool = True) -> None:
    """Compile the checker libraries (building the checker is not using it)."""
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def cpu_has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    fl = set(line.split())
                    return {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd", "fma"} <= fl
    except OSError:
        pass
    return False


def ref_available() -> bool:
    return os.path.exists(REF_SO) and cpu_has_avx512()


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = lib = C.CDLL(ORACLE_SO)
        lib.rgo_distance.restype = C.c_float
        lib.rgo_distance.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
        lib.rgo_distance_batch.restype = None
        lib.rgo_distance_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint64, C.c_void_p]
        lib.rgo_pool_script.restype = C.c_uint32
        lib.rgo_pool_script.argtypes = [C.c_uint32, C.c_uint32] + [C.c_void_p] * 8
        lib.rgo_search_roargraph.restype = C.c_int
        lib.rgo_search_roargraph.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rgo_search_projection_internal.restype = C.c_int
        lib.rgo_search_projection_internal.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                                       C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                                       C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rgo_exact_knn.restype = C.c_int
        lib.rgo_exact_knn.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int,
                                      C.c_uint32, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rgo_compute_recall.restype = C.c_float
        lib.rgo_compute_recall.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.rgo_omp_num_procs.restype = C.c_int

    def num_procs(self) -> int:
        return int(self.lib.rgo_omp_num_procs())

    def distance_batch(self, metric, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        n, d = a.shape
        out = np.empty(n, np.float32)
        self.lib.rgo_distance_batch(metric, _ptr(a), _ptr(b), d, n, _ptr(out))
        return out

    def pool_script(self, capacity, kind, ids, dists):
        return _pool_script(self.lib.rgo_pool_script, capacity, kind, ids, dists)

    def search(self, base, offsets, adj, ep, queries, k, L, metric=1, threads=None):
        base = np.ascontiguousarray(base, np.float32); queries = np.ascontiguousarray(queries, np.float32)
        offsets = np.ascontiguousarray(offsets, np.uint64); adj = np.ascontiguousarray(adj, np.uint32)
        n, d = base.shape
        nq = queries.shape[0]
        assert queries.shape[1] == d
        ids = np.empty((nq, k), np.uint32); dists = np.empty((nq, k), np.float32)
        cmps = np.empty(nq, np.uint32); hops = np.empty(nq, np.uint32)
        sec = C.c_double(0)
        rc = self.lib.rgo_search_roargraph(_ptr(base), n, d, metric, _ptr(offsets), _ptr(adj), ep, _ptr(queries),
                                           nq, k, L, threads or self.num_procs(), _ptr(ids), _ptr(dists),
                                           _ptr(cmps), _ptr(hops), C.byref(sec))
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, seconds=sec.value, rc=rc)

    def search_expanded(self, base, offsets, adj, ep, node_lo, count, L, cap, metric=1, threads=None):
        """Build-time beam searches (SearchProjectionGraphInternal): expanded (ids, dists, count) per target row."""
        base = np.ascontiguousarray(base, np.float32)
        offsets = np.ascontiguousarray(offsets, np.uint64); adj = np.ascontiguousarray(adj, np.uint32)
        n, d = base.shape
        ids = np.zeros((count, cap), np.uint32); dists = np.zeros((count, cap), np.float32)
        cnt = np.zeros(count, np.uint32)
        self.lib.rgo_search_projection_internal(_ptr(base), n, d, metric, _ptr(offsets), _ptr(adj), ep, node_lo, count, L, cap,
                                                threads or self.num_procs(), _ptr(ids), _ptr(dists), _ptr(cnt))
        return ids, dists, cnt

    def exact_knn(self, base, queries, K, metric=1, part_size=0, threads=None):
        base = np.ascontiguousarray(base, np.float32); queries = np.ascontiguousarray(queries, np.float32)
        n, d = base.shape
        nq = queries.shape[0]
        ids = np.empty((nq, K), np.uint32); dists = np.empty((nq, K), np.float32)
        sec = C.c_double(0)
        self.lib.rgo_exact_knn(_ptr(base), n, _ptr(queries), nq, d, metric, K, part_size,
                               threads or self.num_procs(), _ptr(ids), _ptr(dists), C.byref(sec))
        return ids, dists, sec.value

    def recall(self, res, gt, k):
        res = np.ascontiguousarray(res, np.uint32); gt = np.ascontiguousarray(gt, np.uint32)
        return float(self.lib.rgo_compute_recall(res.shape[0], k, gt.shape[1], _ptr(res), _ptr(gt)))


def _pool_script(fn, capacity, kind, ids, dists):
    kind = np.ascontiguousarray(kind, np.uint8); ids = np.ascontiguousarray(ids, np.uint32)
    dists = np.ascontiguousarray(dists, np.float32)
    nops = len(kind)
    oi = np.zeros(capacity + 1, np.uint32); od = np.zeros(capacity + 1, np.float32); of = np.zeros(capacity + 1, np.uint8)
    pop = np.zeros(max(nops, 1), np.uint32); npop = C.c_uint32(0)
    size = fn(capacity, nops, _ptr(kind), _ptr(ids), _ptr(dists), _ptr(oi), _ptr(od), _ptr(of), _ptr(pop),
              C.byref(npop))
    return oi[:size].copy(), od[:size].copy(), of[:size].copy(), pop[:npop.value].copy()


class Ref:
    """The compiled reference.  File based, like the reference's own drivers."""

    def __init__(self):
        if not ref_available():
            raise RuntimeError("reference library unavailable (not built, or host CPU lacks AVX-512)")
        self.lib = lib = C.CDLL(REF_SO)
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_omp_num_procs.restype = C.c_int
        lib.ref_distance_batch.restype = None
        lib.ref_distance_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint, C.c_uint64, C.c_void_p]
        lib.ref_pool_script.restype = C.c_uint32
        lib.ref_pool_script.argtypes = [C.c_uint32, C.c_uint32] + [C.c_void_p] * 8
        lib.ref_build_index.restype = C.c_int
        lib.ref_build_index.argtypes = [C.c_char_p] * 4 + [C.c_int] + [C.c_uint32] * 4 + [C.c_void_p]
        lib.ref_index_open.restype = C.c_void_p
        lib.ref_index_open.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_uint32]
        lib.ref_index_dim.restype = C.c_uint32
        lib.ref_index_dim.argtypes = [C.c_void_p]
        lib.ref_index_search.restype = C.c_int
        lib.ref_index_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_int] + [C.c_void_p] * 5
        lib.ref_index_close.restype = None
        lib.ref_index_close.argtypes = [C.c_void_p]

    def num_procs(self) -> int:
        return int(self.lib.ref_omp_num_procs())

    def err(self) -> str:
        return (self.lib.ref_last_error() or b"").decode()

    def distance_batch(self, metric, a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        n, d = a.shape
        out = np.empty(n, np.float32)
        self.lib.ref_distance_batch(metric, _ptr(a), _ptr(b), d, n, _ptr(out))
        return out

    def pool_script(self, capacity, kind, ids, dists):
        return _pool_script(self.lib.ref_pool_script, capacity, kind, ids, dists)

    def build_index(self, base_fbin, train_fbin, knn_ibin, out_index, metric=1, M_sq=100, M_pjbp=35, L_pjpq=500,
                    threads=1):
        sec = C.c_double(0)
        rc = self.lib.ref_build_index(base_fbin.encode(), train_fbin.encode(), knn_ibin.encode(),
                                      out_index.encode(), metric, M_sq, M_pjbp, L_pjpq, threads, C.byref(sec))
        if rc:
            raise RuntimeError(self.err())
        return sec.value

    def open(self, base_fbin, index_path, metric=1, threads=None):
        h = self.lib.ref_index_open(base_fbin.encode(), index_path.encode(), metric, threads or self.num_procs())
        if not h:
            raise RuntimeError(self.err())
        return h

    def search(self, h, queries, k, L, threads=None, warmup=False):
        from mysteryann_b200.io import pad_rows
        queries = pad_rows(np.asarray(queries, np.float32))
        assert queries.shape[1] == self.lib.ref_index_dim(h)
        nq = queries.shape[0]
        ids = np.zeros((nq, k), np.uint32); dists = np.zeros((nq, k), np.float32)
        cmps = np.zeros(nq, np.uint32); hops = np.zeros(nq, np.uint32)
        sec = C.c_double(0)
        rc = self.lib.ref_index_search(h, _ptr(queries), nq, k, L, threads or self.num_procs(), int(warmup),
                                       _ptr(ids), _ptr(dists), _ptr(cmps), _ptr(hops), C.byref(sec))
        if rc:
            raise RuntimeError(self.err())
        return dict(ids=ids, dists=dists, cmps=cmps, hops=hops, seconds=sec.value, rc=rc)

    def close(self, h):
        self.lib.ref_index_close(h)
