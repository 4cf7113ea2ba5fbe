import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden_mesh(name):
    from fluidity_b200.synthetic import Mesh
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return Mesh(dim=int(z["dim"]), ndglno=np.ascontiguousarray(z["ndglno"], dtype=np.int32),
                X=np.ascontiguousarray(z["X"], dtype=np.float64))


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.lib()
    return oracle


def rel_err(a, ref):
    """Parity metric of SURVEY.md 8(c)(ii): max |a-ref| / max |ref| over the block."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref)) if ref.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - ref)) / scale)


def row_rel_err(a, ref, findrm):
    """Row-wise ||row-row_ref||_inf / ||row_ref||_inf, max over rows (findrm 1-based)."""
    a = np.asarray(a)
    ref = np.asarray(ref)
    starts = (np.asarray(findrm[:-1]) - 1).astype(np.int64)
    num = np.maximum.reduceat(np.abs(a - ref), starts)
    den = np.maximum.reduceat(np.abs(ref), starts)
    ok = den > 0
    out = np.zeros_like(num)
    out[ok] = num[ok] / den[ok]
    out[~ok] = num[~ok]
    return float(out.max())
