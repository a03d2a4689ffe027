"""EfficientConvBlock's second gate in training without a stored input gradient (csrc/stem_tail.cu: pmoe_eca_bn_bwd_sums /
pmoe_eca_bn_bwd_apply; reference basics.py:118-121, conv1 -> BN -> ReLU -> ECA gate -> conv2) through the C-ABI:
 - the per-(image, channel) sums against torch in fp64 on the same bf16 operands;
 - the fused apply against the formula evaluated in fp64 (BatchNorm backward of dz = mask * (dy * gate + dmean)), negative gammas
   and a dead channel included, bf16 storage of the result being the only difference;
 - the stem block's training step with the fusion on and off: every parameter gradient of the two forms agrees to bf16 storage
   precision, and both agree with the CPU oracle."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _operands(n, h, w, c, seed):
    g = torch.Generator().manual_seed(seed)
    raw = (torch.randn(n, h, w, c, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = torch.randn(c, generator=g)
    gamma[0] = 0.0
    beta = torch.randn(c, generator=g) * 0.5
    dy = torch.randn(n, h, w, c, generator=g).to(torch.bfloat16)
    gate = torch.sigmoid(torch.randn(n, c, generator=g))
    dmean = torch.randn(n, c, generator=g) * 0.05
    rd = raw.double()
    mean = rd.mean(dim=(0, 1, 2))
    rstd = 1.0 / torch.sqrt(rd.var(dim=(0, 1, 2), unbiased=False) + 1e-5)
    scale = (gamma.double() * rstd).float()
    shift = (beta.double() - mean * gamma.double() * rstd).float()
    c1 = torch.relu(torch.addcmul(shift, raw.float(), scale)).to(torch.bfloat16)   # what affine_act stores
    return [t.to(dev) for t in (raw, gamma, beta, dy, gate, dmean, mean.float(), rstd.float(), scale, shift, c1)]


@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 7, 9, 16), (1, 1, 3, 8), (5, 16, 16, 128)])
def test_eca_bn_bwd_sums(n, h, w, c):
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    raw, gamma, beta, dy, gate, dmean, mean, rstd, scale, shift, c1 = _operands(n, h, w, c, 3 * n + c)
    p1 = torch.zeros(n, c, dtype=torch.float64, device=dev)
    p2 = torch.zeros(n, c, dtype=torch.float64, device=dev)
    m0 = torch.zeros(n, c, dtype=torch.float64, device=dev)
    vd, vc = view4(dy), view4(c1)
    check(lib().pmoe_eca_bn_bwd_sums(C.byref(vd), C.byref(vc), p1.data_ptr(), p2.data_ptr(), m0.data_ptr(), c, stream_ptr()), "sums")
    on = (c1 > 0).double()
    assert torch.equal(m0, on.sum(dim=(1, 2)))
    ref1 = (dy.double() * on).sum(dim=(1, 2))
    ref2 = (dy.double() * c1.double()).sum(dim=(1, 2))
    tol1 = 1e-5 * (dy.double() * on).abs().sum(dim=(1, 2)) + 1e-9     # fp32 partial sums per block, fp64 across blocks
    tol2 = 1e-5 * (dy.double() * c1.double()).abs().sum(dim=(1, 2)) + 1e-9
    assert bool(((p1 - ref1).abs() <= tol1).all()) and bool(((p2 - ref2).abs() <= tol2).all())


@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 7, 9, 16), (1, 1, 3, 8), (5, 16, 16, 128)])
def test_eca_bn_bwd_apply_vs_fp64(n, h, w, c):
    from pmoe_b200 import train
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    raw, gamma, beta, dy, gate, dmean, mean, rstd, scale, shift, c1 = _operands(n, h, w, c, 7 * n + c)
    # the mask exactly as the forward computed it: fp32 fma of the bf16 input with the fp32 affine
    m = (torch.addcmul(shift, raw.float(), scale) > 0).double()
    dz = m * (dy.double() * gate.double()[:, None, None, :] + dmean.double()[:, None, None, :])
    xhat = (raw.double() - mean.double()) * rstd.double()
    s1 = dz.sum(dim=(0, 1, 2))
    s2 = (dz * xhat).sum(dim=(0, 1, 2))
    N = n * h * w
    ref = gamma.double() * rstd.double() * (dz - s1 / N - xhat * s2 / N)
    dx = torch.empty_like(raw)
    dgam = torch.full((c,), 7.0, device=dev)
    dbet = torch.full((c,), -3.0, device=dev)
    pg = train.BnParamGrads()
    pg.dgamma, pg.dbeta, pg.n, pg.accumulate = dgam.data_ptr(), dbet.data_ptr(), c, 1
    vd, vr, vx = view4(dy), view4(raw), view4(dx)
    check(lib().pmoe_eca_bn_bwd_apply(C.byref(vd), C.byref(vr), gate.data_ptr(), gate.stride(0), dmean.data_ptr(), dmean.stride(0),
                                      scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                      s1.data_ptr(), s2.data_ptr(), 1.0 / N, C.byref(vx), C.byref(pg), stream_ptr()), "apply")
    assert _rel(dx.float(), ref) < 4e-3                       # bf16 storage of dx
    assert _rel(dgam.double() - 7.0, s2) < 1e-5 and _rel(dbet.double() + 3.0, s1) < 1e-5


def test_eca_bn_bwd_rejects_strided():
    from pmoe_b200._lib import lib, view4, stream_ptr
    a = torch.zeros(1, 4, 4, 32, dtype=torch.bfloat16, device=dev)[..., :16]      # channel slice: not dense
    b = torch.zeros(1, 4, 4, 16, dtype=torch.bfloat16, device=dev)
    z = torch.zeros(1, 16, dtype=torch.float64, device=dev)
    va, vb = view4(a), view4(b)
    assert lib().pmoe_eca_bn_bwd_sums(C.byref(va), C.byref(vb), z.data_ptr(), z.data_ptr(), z.data_ptr(), 16, stream_ptr()) == -2


@pytest.mark.parametrize("cin,cout", [(12, 64), (24, 32)])
def test_stem_block_step_with_and_without_lazy_gate_gradient(cin, cout):
    from pmoe_b200 import config, train
    from pmoe_b200.model.blocks.basics import EfficientConvBlock
    from oracle import functional as O
    torch.manual_seed(11)
    x = torch.rand(4, cin, 32, 48, device=dev)
    cot = torch.randn(4, cout, 32, 48, device=dev)
    grads, outs = {}, {}
    with config.use_precision("bf16"):
        blk = EfficientConvBlock(cin, cout).cuda().train()
        sd = {k: v.clone() for k, v in blk.state_dict().items()}
        for fused in (True, False):
            blk.load_state_dict(sd)
            blk.zero_grad(set_to_none=True)
            old = train.FUSE_ECA_BN_BWD
            train.FUSE_ECA_BN_BWD = fused
            try:
                y = blk(x)
                (y * cot).sum().backward()
            finally:
                train.FUSE_ECA_BN_BWD = old
            outs[fused] = y.detach().clone()
            grads[fused] = {k: p.grad.detach().clone() for k, p in blk.named_parameters()}
    assert torch.equal(outs[True], outs[False]) or _rel(outs[True], outs[False]) < 1e-3   # the forward is the same code
    num = sum((grads[True][k].double() - g.double()).pow(2).sum() for k, g in grads[False].items())
    den = sum(g.double().pow(2).sum() for g in grads[False].values())
    print("\n   lazy vs stored gate gradient: all parameters %.3e, worst parameter %.3e"
          % ((num / den).sqrt().item(), max(_rel(grads[True][k], g) for k, g in grads[False].items())))
    assert (num / den).sqrt().item() < 1e-2
    for k, g in grads[False].items():   # the ECA conv1d weights (3 values each) are cancelling sums: more slack per parameter
        assert _rel(grads[True][k], g) < 3e-2, (k, _rel(grads[True][k], g))
    # both against the fp32 CPU oracle on the same weights: the form that never rounds the gate's input gradient to bf16 is not
    # further from it than the form that stores it (bf16 arithmetic of the product either way)
    sdg = {k: (v.detach().cpu().clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))
               else v.detach().cpu().clone()) for k, v in sd.items()}
    ref = O.eca_conv_block(x.cpu(), sdg, "", True)
    (ref * cot.cpu()).sum().backward()
    err = {}
    for fused in (True, False):
        num = sum((grads[fused][k].cpu().double() - sdg[k].grad.double()).pow(2).sum() for k in grads[fused])
        den = sum(sdg[k].grad.double().pow(2).sum() for k in grads[fused])
        err[fused] = (num / den).sqrt().item()
    print("\n   stem block gradients vs the fp32 oracle: lazy %.3e | stored %.3e" % (err[True], err[False]))
    assert err[True] < 1.5 * err[False] + 2e-3, err


def test_relu_threshold_is_exact_for_every_bf16_input():
    """st_relu_threshold (bisection per channel): the packed compare x' > T must reproduce fmaf(x, scale, shift) > 0 for EVERY finite
    bf16 x. With gamma = rstd = gate = 1, dmean = 0 and zero backward sums, pmoe_eca_bn_bwd_apply returns dx = mask * dy."""
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    c = 64
    bits = torch.arange(65536, dtype=torch.int32)
    xall = bits.to(torch.int16).view(torch.bfloat16)                       # every bf16 pattern once
    finite = torch.isfinite(xall.float())
    x = xall.view(1, 256, 256, 1).expand(1, 256, 256, c).contiguous().to(dev)
    g = torch.Generator().manual_seed(17)
    scale = torch.randn(c, generator=g) * torch.tensor([10.0 ** ((i % 9) - 4) for i in range(c)])
    shift = torch.randn(c, generator=g) * torch.tensor([10.0 ** (((i // 3) % 9) - 4) for i in range(c)])
    scale[0], shift[0] = 0.0, 1.0
    scale[1], shift[1] = 0.0, -1.0
    scale[2], shift[2] = 1.0, 0.0
    scale[3], shift[3] = -1.0, 0.0
    scale[4], shift[4] = 3e-39, 1e-38                                      # denormal scale
    scale, shift = scale.to(dev), shift.to(dev)
    dy = torch.ones_like(x)
    ones = torch.ones(1, c, device=dev)
    zeros = torch.zeros(1, c, device=dev)
    one_c, zero_c = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    z64 = torch.zeros(c, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    vd, vr, vx = view4(dy), view4(x), view4(dx)
    check(lib().pmoe_eca_bn_bwd_apply(C.byref(vd), C.byref(vr), ones.data_ptr(), ones.stride(0), zeros.data_ptr(), zeros.stride(0),
                                      scale.data_ptr(), shift.data_ptr(), zero_c.data_ptr(), one_c.data_ptr(), one_c.data_ptr(),
                                      z64.data_ptr(), z64.data_ptr(), 1.0 / 65536, C.byref(vx), None, stream_ptr()), "apply")
    got = dx.view(65536, c).float() > 0.5
    want = torch.addcmul(shift, x.view(65536, c).float(), scale) > 0          # one fused multiply-add in fp32, as the forward kernels do
    fin = finite.to(dev)
    bad = (got != want)[fin]
    assert not bool(bad.any()), (int(bad.sum()), torch.nonzero(bad)[:8].tolist())
