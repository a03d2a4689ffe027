"""GPU parity of the tensor-core weight-gradient kernels (halo + streaming variants) against the CUDA-core kernel on
the same bf16 operands (identical products, fp32 accumulation in a different order: 1e-4) and against torch's fp32
conv weight gradient on small cases."""
import importlib.util
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_wgrad_tc_matches_simt_and_torch():
    spec = importlib.util.spec_from_file_location("gpu_check_wgrad", os.path.join(ROOT, "scripts", "gpu_check_wgrad.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.run("64->64 32x32 B2", 2, 32, 32, [64], 64)
    m.run("64->64 48x40 B1 (partial)", 1, 48, 40, [64], 64)
    m.run("128->128 32x32 B2", 2, 32, 32, [128], 128)
    m.run("64+64->64 32x32 B2 (concat)", 2, 32, 32, [64, 64], 64)
    m.run("256->256 32x32 B1", 1, 32, 32, [256], 256)
    m.run("512->512 16x16 B2", 2, 16, 16, [512], 512)
    m.general_cases(False)
    done = [r for r in m.report if r.get("supported")]
    assert len(done) >= 12, [r["name"] for r in m.report]
    for r in done:
        assert r["rel_vs_simt"] < 1e-4, r
        if r.get("rel_vs_torch") is not None:
            assert r["rel_vs_torch"] < 1e-4, r
