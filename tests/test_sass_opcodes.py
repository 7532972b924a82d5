"""CPU test of the built library's SASS (cuobjdump, no GPU needed): every hot kernel takes the Blackwell path its DESIGN.md
section claims, and the packed-FP32 distance loop keeps the product and the sum separately rounded.

The second point is a parity guard: the reference's main loop is an unfused vmulps + vaddps (include/efanna2e/distance.h:39-89,
179-223 under g++ -Ofast), and ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2.  rg_distance.cuh therefore issues the
product as fma(a, b, -0.0) with an addend the compiler cannot see.  Were a toolchain ever to fold that, the FADD2 of the main
loop would disappear from the SASS - which is what this test looks for."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def kernels():
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    from mysteryann_b200 import build

    build.build()
    import sass_opcodes

    per = sass_opcodes.histogram(os.path.join(ROOT, "mysteryann_b200", "libroargraph_b200.so"))
    names = list(per)
    pretty = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    return {p: per[m] for m, p in zip(names, pretty)}, sass_opcodes.family


def test_search_kernels_use_tma_bulk_and_packed_fp32(kernels):
    per, family = kernels
    k1 = {n: c for n, c in per.items() if "rg_search_kernel<" in n}
    assert len(k1) >= 16, sorted(k1)
    for name, c in k1.items():
        # (bool kIP, int kGather, ...): gather 2 = one TMA bulk copy per row, 1 = cp.async
        gather = int(name.split("rg_search_kernel<")[1].split(",")[1].strip().replace("(int)", ""))
        if gather == 2:
            assert family(c, "UBLKCP") >= 1, name
        else:
            assert family(c, "LDGSTS") >= 1, name
        ffma2, fadd2 = family(c, "FFMA2"), family(c, "FADD2")
        assert ffma2 > 0 and family(c, "FMUL") == 0, (name, ffma2, family(c, "FMUL"))
        ip = "<(bool)1" in name
        # inner product: one FADD2 per FFMA2 (acc + RN(v*q)); L2: two (v - q, then acc + RN(d*d))
        assert fadd2 == (1 if ip else 2) * ffma2, (name, ffma2, fadd2)


def test_prune_kernels_use_packed_fp32(kernels):
    per, family = kernels
    prune = {n: c for n, c in per.items() if "prune_kernel<" in n}
    assert len(prune) == 2, sorted(prune)
    for name, c in prune.items():
        ffma2, fadd2 = family(c, "FFMA2"), family(c, "FADD2")
        assert ffma2 > 0 and family(c, "FMUL") == 0, name
        assert fadd2 == (1 if "<(bool)1" in name else 2) * ffma2, (name, ffma2, fadd2)


def test_knn_gemm_is_tcgen05_tmem_tma(kernels):
    per, family = kernels
    gemm = {n: c for n, c in per.items() if "knn_gemm_filter_kernel<" in n}
    assert gemm, sorted(per)
    for name, c in gemm.items():
        assert family(c, "UTCHMMA") >= 1, name     # tcgen05.mma kind::f16, cta_group::2
        assert family(c, "LDTM") >= 1, name        # tcgen05.ld: accumulators read back from TMEM
        assert family(c, "UTMALDG") >= 1, name     # TMA tensor loads of the operand tiles
        assert family(c, "HMMA") == 0, name        # no legacy mma.sync path
