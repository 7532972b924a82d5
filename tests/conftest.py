import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ("ip_d200", "l2_d48", "ip_d24_norm")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_case(name):
    """Golden case -> dict with fp32 arrays, CSR graph and per-L reference outputs."""
    from mysteryann_b200 import io

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = {k: z[k] for k in z.files}
    for k in ("base", "train", "test"):
        c[k] = c[k].astype(np.float32)
    c["metric"], c["M_sq"], c["M_pjbp"], c["L_pjpq"] = (int(v) for v in c["params"])
    raw = c["index"].view(np.uint32)
    ep, off, adj = io.parse_index(raw)
    c["ep"], c["offsets"], c["adj"] = ep, off, adj
    return c


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.binding import Ref, ref_available

    if not ref_available():
        pytest.skip("oracle/_ref not built or host CPU lacks AVX-512")
    return Ref()
