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


@pytest.mark.parametrize("cout", [1, 4])
def test_skinny_linear_wgrad_uses_padded_tensor_core_path(cout):
    """The 512->1 / 512->4 expert-head Linears (model/moe.py:71-72) over many rows: the weight gradient of a layer with fewer than
    64 output channels runs on the tensor-core kernel through a zero-padded dy (ops.conv_wgrad) and must equal torch's dy^T x on
    the same bf16 operands (fp32 accumulation: 1e-4), also when accumulated into a non-zero dwpack."""
    import torch
    from pmoe_b200 import ops, profiler
    rows, cin = 16384, 512
    g = torch.Generator().manual_seed(cout)
    x = torch.randn(1, 1, rows, cin, generator=g).cuda().to(torch.bfloat16)
    dy = torch.zeros(1, 1, rows, 16, dtype=torch.bfloat16, device="cuda")
    dy[..., :cout] = torch.randn(1, 1, rows, cout, generator=g).cuda().to(torch.bfloat16)
    segs = ops.conv_segments([(0, 0)], [cin], 64)
    base = torch.randn(16, cin, generator=g).cuda()
    dwp = base.clone()
    profiler.reset()
    profiler.enable_events(True)
    ops.conv_wgrad([x], segs, 64, dy, dwp, tag="skinny")
    kinds = [r[0] for r in profiler.records()]
    profiler.enable_events(False)
    assert "conv_wgrad_tc" in kinds and "conv_wgrad" not in kinds, kinds
    ref = dy[0, 0].float().t() @ x[0, 0].float() + base
    assert ((dwp - ref).norm() / ref.norm()).item() < 1e-4
    assert (dwp[cout:] - base[cout:]).abs().max().item() == 0
