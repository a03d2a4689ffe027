import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def bf16_eval_errors(out, ref_fp32, ref_bf16_storage):
    """(distance to the fp32 oracle, distance to the oracle run with bf16 storage, that oracle's own distance to fp32)."""
    return rel_err(out, ref_fp32), rel_err(out, ref_bf16_storage), rel_err(ref_bf16_storage, ref_fp32)


def assert_bf16_eval(out, ref_fp32, ref_bf16_storage, what=""):
    """north_star bf16 bound for outputs: within 1e-2 of the reference in bf16 (the oracle evaluated with the product's storage
    points rounded to bf16) and, against the fp32 reference, never further than 1e-2 or 1.25 x what that bf16 evaluation itself
    costs (small random-init geometries put ~1e-2 of storage noise on a 23-layer U-Net, chains of U-Nets more)."""
    e32, eemu, eself = bf16_eval_errors(out, ref_fp32, ref_bf16_storage)
    print("[bf16 eval %s] vs fp32 oracle %.3e | vs bf16-storage oracle %.3e | bf16-storage oracle vs fp32 %.3e" % (what, e32, eemu, eself))
    # two bf16-storage evaluations whose rounding noise is independent are sqrt(2) x that noise apart: where the evaluation itself
    # loses more than 1e-2 against fp32 (chains of U-Nets), the distance to it is bounded by 1.5 x its own distance to fp32
    assert eemu < max(1e-2, 1.5 * eself), (what, eemu, eself)
    assert e32 < max(1e-2, 1.25 * eself), (what, e32, eself)
    return e32, eemu, eself
