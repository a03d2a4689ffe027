"""Max-pool forward (+ argmax codes) and backward through the C-ABI against torch on the same bf16 operands: the pooled
values and the routing of every gradient are exact (ties go to the first maximum in row-major window order, as ATen
does); the backward sum of up to four bf16 gradients is compared after bf16 rounding with a one-ulp allowance.
Covers the specialised 3x3/s2/p1 and 2x2/s2 kernels, the accumulate form and the generic fallback (odd geometry)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def _run(n, h, w, c, k, s, p, accumulate, gen):
    from pmoe_b200 import nhwc
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    from pmoe_b200.nhwc import Act, dtype_code
    # few distinct values -> many ties inside a window
    x = (torch.randint(-6, 7, (n, c, h, w), generator=gen).float() / 4).to(dev)
    xt = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    y, idx = nhwc.maxpool(Act(xt, c), k, s, p, want_idx=True)
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, k, s, p)
    assert torch.equal(y.t.float().permute(0, 3, 1, 2), yr.detach())
    dy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(5)).to(dev).to(torch.bfloat16)
    yr.backward(dy.float())
    dyt = dy.permute(0, 2, 3, 1).contiguous()
    base = torch.randn(n, h, w, c, generator=torch.Generator().manual_seed(6)).to(dev).to(torch.bfloat16)
    dx = base.clone() if accumulate else torch.full((n, h, w, c), 7.0, dtype=torch.bfloat16, device=dev)
    vdy, vdx = view4(dyt), view4(dx)
    check(lib().pmoe_maxpool_bwd_idx(C.byref(vdy), idx.data_ptr(), C.byref(vdx), dtype_code(dx), k, s, p, int(accumulate),
                                     stream_ptr()), "maxpool_bwd_idx")
    ref = xr.grad.permute(0, 2, 3, 1)
    if accumulate:
        ref = ref + base.float()
    got = dx.float()
    ulp = ref.abs().clamp_min(1e-3) * 2.0 ** -7
    assert ((got - ref).abs() <= ulp).all()
    # gradient routing itself is exact: same zero pattern when nothing is accumulated
    if not accumulate:
        assert torch.equal(got != 0, ref.to(torch.bfloat16).float() != 0)


@pytest.mark.parametrize("k,s,p", [(3, 2, 1), (2, 2, 0)])
@pytest.mark.parametrize("accumulate", [False, True])
def test_maxpool_fast_kernels(k, s, p, accumulate):
    gen = torch.Generator().manual_seed(17)
    _run(3, 22, 18, 64, k, s, p, accumulate, gen)
    _run(2, 9, 11, 48, k, s, p, accumulate, gen)   # odd sizes, channel-group count not a power of two


def test_maxpool_generic_geometry():
    gen = torch.Generator().manual_seed(18)
    _run(2, 13, 10, 16, 3, 1, 1, False, gen)
