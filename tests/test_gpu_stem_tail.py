"""The fused tail of the ResNet stem (csrc/stem_tail.cu: torchvision's bn1 -> relu -> MaxPool2d(3, 2, 1) behind
conv1 := EfficientConvBlock, reference backbone.py:57-61) through the C-ABI:
 - forward against the separate product kernels (bit-exact: the affine + ReLU is monotone, so pooling the raw tensor and
   transforming the maximum stores the same bf16 values) and against torch in fp64 on the same bf16 operands;
 - backward (dx, d gamma, d beta and the upstream BatchNorm's two sums) against torch autograd in fp64 through the whole
   BatchNorm(batch statistics) -> ReLU -> max_pool2d chain, negative gammas included;
 - a ResNet-18 training step with the fusion on and off."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _operands(n, h, w, c, seed, dead_channel=True):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, h, w, c, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = torch.randn(c, generator=g)            # both signs
    if dead_channel:
        gamma[0] = 0.0                              # a dead channel: relu(beta) everywhere
    beta = torch.randn(c, generator=g) * 0.5
    dp = torch.randn(n, h // 2, w // 2, c, generator=g).to(torch.bfloat16)
    xd = x.double()
    mean = xd.mean(dim=(0, 1, 2))
    var = xd.var(dim=(0, 1, 2), unbiased=False)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    scale = (gamma.double() * rstd).float()
    shift = (beta.double() - mean * gamma.double() * rstd).float()
    return x, gamma, beta, dp, mean.float(), rstd.float(), scale, shift


def _fused_forward(x, scale, shift):
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    n, h, w, c = x.shape
    p = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=dev)
    xm = torch.empty_like(p)
    idx = torch.empty(p.shape, dtype=torch.uint8, device=dev)
    vx, vp = view4(x), view4(p)
    check(lib().pmoe_bn_relu_maxpool_fwd(C.byref(vx), scale.data_ptr(), shift.data_ptr(), C.byref(vp), idx.data_ptr(), xm.data_ptr(),
                                         stream_ptr()), "bn_relu_maxpool_fwd")
    return p, idx, xm


@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 8, 6, 16), (1, 2, 2, 8), (2, 32, 32, 128)])
def test_stem_tail_forward(n, h, w, c):
    from pmoe_b200 import nhwc
    x, gamma, beta, dp, mean, rstd, scale, shift = [t.to(dev) for t in _operands(n, h, w, c, 11 * n + c)]
    p, idx, xm = _fused_forward(x, scale, shift)
    z = nhwc.affine_act(x, scale, shift, "relu")
    ref = nhwc.maxpool(nhwc.Act(z, c), 3, 2, 1).t
    assert torch.equal(p, ref)
    z64 = torch.relu(x.double() * scale.double() + shift.double()).permute(0, 3, 1, 2)
    p64, i64 = torch.nn.functional.max_pool2d(z64, 3, 2, 1, return_indices=True)
    assert _rel(p.float(), p64.permute(0, 2, 3, 1)) < 4e-3
    # x at the argmax reproduces the pooled value, and is a member of the window the code points to
    assert torch.equal(torch.relu(torch.addcmul(shift, xm.float(), scale)).to(torch.bfloat16), p) or \
        _rel(torch.relu(xm.float() * scale + shift), p.float()) < 4e-3
    r, cc = (idx // 3).long(), (idx % 3).long()
    oh = torch.arange(h // 2, device=dev).view(1, -1, 1, 1)
    ow = torch.arange(w // 2, device=dev).view(1, 1, -1, 1)
    ih, iw = oh * 2 - 1 + r, ow * 2 - 1 + cc
    assert ih.min().item() >= 0 and ih.max().item() < h and iw.min().item() >= 0 and iw.max().item() < w
    ni = torch.arange(n, device=dev).view(-1, 1, 1, 1).expand_as(ih)
    ci = torch.arange(c, device=dev).view(1, 1, 1, -1).expand_as(ih)
    assert torch.equal(x[ni, ih, iw, ci], xm)


@pytest.mark.parametrize("chain", [False, True])
@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 8, 6, 16), (1, 2, 2, 8), (2, 32, 32, 128)])
def test_stem_tail_backward_vs_autograd(n, h, w, c, chain):
    from pmoe_b200 import train
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    # (a channel whose gamma is EXACTLY 0 has a constant output: every window is a tie, torch routes to its first element, the
    # fused kernel to its largest input; dx and d beta do not depend on that choice, d gamma of that one channel does)
    x, gamma, beta, dp, mean, rstd, scale, shift = [t.to(dev) for t in _operands(n, h, w, c, 5 * n + c, dead_channel=False)]
    p, idx, xm = _fused_forward(x, scale, shift)
    s1 = torch.zeros(c, dtype=torch.float64, device=dev)
    s2 = torch.zeros(c, dtype=torch.float64, device=dev)
    vdp, vx = view4(dp), view4(x)
    check(lib().pmoe_bn_relu_maxpool_bwd_reduce(C.byref(vdp), xm.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                                rstd.data_ptr(), s1.data_ptr(), s2.data_ptr(), None, stream_ptr()), "reduce")
    dx = torch.empty_like(x)
    vdx = view4(dx)
    n1 = torch.zeros(c, dtype=torch.float64, device=dev) if chain else None
    n2 = torch.zeros(c, dtype=torch.float64, device=dev) if chain else None
    dgam = torch.full((c,), 7.0, device=dev)
    dbet = torch.full((c,), -3.0, device=dev)
    pg = train.BnParamGrads()
    pg.dgamma, pg.dbeta, pg.n, pg.accumulate = dgam.data_ptr(), dbet.data_ptr(), c, 1 if chain else 0
    check(lib().pmoe_bn_relu_maxpool_bwd_apply(
        C.byref(vdp), idx.data_ptr(), C.byref(vx), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
        s1.data_ptr(), s2.data_ptr(), 1.0 / (n * h * w), C.byref(vdx), None if n1 is None else n1.data_ptr(),
        None if n2 is None else n2.data_ptr(), C.byref(pg), stream_ptr()), "apply")
    # fp64 autograd through batch-statistics BatchNorm -> ReLU -> max pool on the same bf16 operands
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    mu = xd.mean(dim=(0, 1, 2))
    var = xd.var(dim=(0, 1, 2), unbiased=False)
    z = torch.relu((xd - mu) / torch.sqrt(var + 1e-5) * gd + bd)
    pp = torch.nn.functional.max_pool2d(z.permute(0, 3, 1, 2), 3, 2, 1)
    pp.backward(dp.double().permute(0, 3, 1, 2))
    assert _rel(dx.float(), xd.grad) < 4e-3                       # bf16 storage of dx
    off = 7.0 if chain else 0.0
    assert _rel(dgam.double() - off, gd.grad) < 1e-4 and _rel(dbet.double() - (-3.0 if chain else 0.0), bd.grad) < 1e-4
    if chain:   # the upstream BatchNorm's backward sums, taken from the fp32 dx before it is rounded for storage: they differ from
        # the sums of the STORED dx (what a separate reduce pass reads back) by a zero-mean sum of bf16 rounding errors,
        # |err_i| <= 2^-9 |dx_i|, i.e. a few 2^-9 * sqrt(sum dx^2) per channel
        dxd = dx.double()
        tol1 = 4 * 2.0 ** -9 * (dxd * (x > 0)).pow(2).sum(dim=(0, 1, 2)).sqrt() + 1e-6 * dxd.abs().sum(dim=(0, 1, 2))
        assert bool(((n1 - (dxd * (x > 0)).sum(dim=(0, 1, 2))).abs() <= tol1).all())
        tol2 = 4 * 2.0 ** -9 * (dxd * x.double()).pow(2).sum(dim=(0, 1, 2)).sqrt() + 1e-6 * (dxd * x.double()).abs().sum(dim=(0, 1, 2))
        assert bool(((n2 - (dxd * x.double()).sum(dim=(0, 1, 2))).abs() <= tol2).all())


def test_stem_tail_rejects_other_geometries():
    from pmoe_b200._lib import lib, view4, stream_ptr
    x = torch.zeros(1, 5, 6, 16, dtype=torch.bfloat16, device=dev)     # odd height
    p = torch.zeros(1, 2, 3, 16, dtype=torch.bfloat16, device=dev)
    sc = torch.ones(16, device=dev)
    idx = torch.zeros(p.shape, dtype=torch.uint8, device=dev)
    vx, vp = view4(x), view4(p)
    assert lib().pmoe_bn_relu_maxpool_fwd(C.byref(vx), sc.data_ptr(), sc.data_ptr(), C.byref(vp), idx.data_ptr(), p.data_ptr(),
                                          stream_ptr()) != 0


def test_resnet18_step_with_and_without_fused_stem_tail():
    """A ResNet-18 training step with the fusion on and off. The bf16 step is not run-to-run reproducible (the order of the
    atomics in the BatchNorm statistics flips bf16 roundings, and batch-statistics BatchNorm over 4 x 2 x 2 values in layer4
    amplifies them: measured A/A 4e-3 on the features, 0.2 on the gradients, scripts/gpu_stem_tail_diag.py), so the A/B
    difference is held to the measured A/A floor of the same step; what CAN be bit-exact is: the pooled tensor the fused op
    returns equals the separate launches' on the same stem output."""
    from pmoe_b200 import config, train
    from pmoe_b200.model.blocks.backbone import get_backbone
    torch.manual_seed(3)
    x = torch.rand(4, 12, 64, 64, device=dev)
    cot = torch.randn(4, 512, device=dev)
    runs = {True: [], False: []}
    seen = []
    orig = train.bn_relu_maxpool_op

    def capture(tape, bn, stem, tag=""):
        pa = orig(tape, bn, stem, tag=tag)
        old = train.FUSE_STEM_TAIL
        train.FUSE_STEM_TAIL = not old
        save, tape.save = tape.save, False
        try:   # the other form on the same stem output (forward only; running statistics are restored by load_state_dict)
            other = orig(tape, bn, stem, tag=tag)
        finally:
            train.FUSE_STEM_TAIL, tape.save = old, save
        seen.append(torch.equal(pa.t, other.t))
        return pa
    old = train.FUSE_STEM_TAIL
    train.bn_relu_maxpool_op = capture
    try:
        with config.use_precision("bf16"):
            net = get_backbone(arch="resnet18", n_frames=4, pretrained=False, gamma=2, b=1, n_channels=3).cuda().train()
            sd = {k: v.clone() for k, v in net.state_dict().items()}
            for fused in (True, False, True, False):
                net.load_state_dict(sd)
                net.zero_grad(set_to_none=True)
                train.FUSE_STEM_TAIL = fused
                f = net(x)
                (f * cot).sum().backward()
                runs[fused].append((f.detach().clone(), {n: p.grad.detach().clone() for n, p in net.named_parameters()}))
    finally:
        train.FUSE_STEM_TAIL = old
        train.bn_relu_maxpool_op = orig
    assert seen and all(seen)

    def gdist(a, b):
        num = sum((a[n].double() - g.double()).pow(2).sum() for n, g in b.items())
        den = sum(g.double().pow(2).sum() for g in b.values())
        return (num / den).sqrt().item()
    floor_f = max(_rel(runs[True][1][0], runs[True][0][0]), _rel(runs[False][1][0], runs[False][0][0]))
    floor_g = max(gdist(runs[True][1][1], runs[True][0][1]), gdist(runs[False][1][1], runs[False][0][1]))
    ab_f = min(_rel(a[0], b[0]) for a in runs[True] for b in runs[False])
    ab_g = min(gdist(a[1], b[1]) for a in runs[True] for b in runs[False])
    # (two runs of one form are sometimes bit-identical — then the floor of that pair is 0 while the other form still rounds the
    # gradient differently at a few places, which the 16-value BatchNorms of layer4 amplify: the additive terms cover that case;
    # exactness of the kernels themselves is pinned by the tests above)
    assert ab_f < 2 * floor_f + 1e-2, (ab_f, floor_f)
    assert ab_g < 2 * floor_g + 0.6, (ab_g, floor_g)   # (A/A itself measured up to 0.32; unrelated gradients would give ~1.4)


def test_stem_backward_kernels_at_bench_size_properties():
    """BASELINE configs[2] geometry (one micro-batch of the 224^2 x 64-channel stem tensors; 128 images here to bound memory):
    size-independent properties of a BatchNorm backward — its output is orthogonal to the constant and to x-hat per channel
    (sum dx = 0, sum dx * xhat = 0 up to the bf16 storage of dx) — for the fused stem tail and for the fused ECA + BatchNorm apply,
    and the fused forward equals the separate launches bit for bit."""
    from pmoe_b200 import nhwc, train
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    n, h, w, c = 128, 224, 224, 64
    g = torch.Generator(device=dev).manual_seed(5)
    x = (torch.randn(n, h, w, c, generator=g, device=dev) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(c, generator=g, device=dev) + 0.5
    beta = torch.randn(c, generator=g, device=dev) * 0.3
    s_ = torch.zeros(c, dtype=torch.float64, device=dev)
    q_ = torch.zeros(c, dtype=torch.float64, device=dev)
    vx = view4(x)
    check(lib().pmoe_channel_stats(C.byref(vx), 1, s_.data_ptr(), q_.data_ptr(), stream_ptr()), "stats")
    N = n * h * w
    mean64 = s_ / N
    rstd64 = 1.0 / torch.sqrt((q_ / N - mean64 * mean64).clamp_min(0) + 1e-5)
    scale = (gamma.double() * rstd64).float()
    shift = (beta.double() - mean64 * gamma.double() * rstd64).float()
    mean, rstd = mean64.float(), rstd64.float()
    # ---- stem tail
    p, idx, xm = _fused_forward(x, scale, shift)
    z = nhwc.affine_act(x, scale, shift, "relu")
    assert torch.equal(p, nhwc.maxpool(nhwc.Act(z, c), 3, 2, 1).t)
    del z
    dp = (torch.randn(n, h // 2, w // 2, c, generator=g, device=dev) + 0.7).to(torch.bfloat16)   # non-zero mean: sum dz ~ sum |dz|
    s1 = torch.zeros(c, dtype=torch.float64, device=dev)
    s2 = torch.zeros(c, dtype=torch.float64, device=dev)
    vdp = view4(dp)
    check(lib().pmoe_bn_relu_maxpool_bwd_reduce(C.byref(vdp), xm.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                                rstd.data_ptr(), s1.data_ptr(), s2.data_ptr(), None, stream_ptr()), "reduce")
    dx = torch.empty_like(x)
    vdx = view4(dx)
    check(lib().pmoe_bn_relu_maxpool_bwd_apply(
        C.byref(vdp), idx.data_ptr(), C.byref(vx), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
        s1.data_ptr(), s2.data_ptr(), 1.0 / N, C.byref(vdx), None, None, None, stream_ptr()), "apply")

    def orthogonality(dx_):
        a = torch.zeros(c, dtype=torch.float64, device=dev)
        b = torch.zeros(c, dtype=torch.float64, device=dev)
        ab = torch.zeros(c, dtype=torch.float64, device=dev)
        for i in range(0, n, 16):   # chunks bound the fp64 temporaries
            d = dx_[i:i + 16].double()
            xh = (x[i:i + 16].double() - mean64) * rstd64
            a += d.sum(dim=(0, 1, 2))
            b += (d * xh).sum(dim=(0, 1, 2))
            ab += d.abs().sum(dim=(0, 1, 2))
        return (a.abs() / ab).max().item(), (b.abs() / ab).max().item()
    o1, o2 = orthogonality(dx)
    # Without the two subtracted means the sums would be ~0.5 of sum|dx| (the pooled gradient has mean 0.7). What is left is the bf16
    # storage of dx, and it is NOT sqrt(N)-small: the routed values are A[c] * (a bf16 number), a few hundred distinct products per
    # channel whose rounding errors repeat — rounding the exact fp64 result to bf16 leaves 1e-4 of sum|dx|
    # (scripts/gpu_stem_tail_sumcheck.py: 0.9e-4 for the rounded exact result, 1.7e-4 for the kernel).
    print("\n   stem tail at %dx%dx%dx%d: |sum dx| / sum|dx| = %.2e, |sum dx xhat| / sum|dx| = %.2e" % (n, h, w, c, o1, o2))
    assert o1 < 1e-3 and o2 < 1e-3, (o1, o2)
    # ---- ECA + BatchNorm apply on the same operands (dy at full resolution)
    del dx, dp, p, idx, xm
    dy = (torch.randn(n, h, w, c, generator=g, device=dev) + 0.7).to(torch.bfloat16)
    gate = torch.sigmoid(torch.randn(n, c, generator=g, device=dev))
    dmean = torch.randn(n, c, generator=g, device=dev) * 0.05
    c1 = nhwc.affine_act(x, scale, shift, "relu")
    p1 = torch.zeros(n, c, dtype=torch.float64, device=dev)
    p2 = torch.zeros(n, c, dtype=torch.float64, device=dev)
    m0 = torch.zeros(n, c, dtype=torch.float64, device=dev)
    vdy, vc1 = view4(dy), view4(c1)
    check(lib().pmoe_eca_bn_bwd_sums(C.byref(vdy), C.byref(vc1), p1.data_ptr(), p2.data_ptr(), m0.data_ptr(), c, stream_ptr()), "sums")
    pool = nhwc.channel_sums(c1)[:, :c].double()
    t1 = (gate.double() * p1 + dmean.double() * m0).sum(0)                 # sum dc1 * [c1 > 0]
    t2 = (gate.double() * p2 + dmean.double() * pool).sum(0)               # sum dc1 * c1
    sc, sh = scale.double(), shift.double()
    sraw = (t2 - sh * t1) / sc                                             # sum dc1 * m * raw through the forward affine
    b1, b2 = t1.contiguous(), (rstd64 * (sraw - mean64 * t1)).contiguous()
    del c1
    draw = torch.empty_like(x)
    vdr = view4(draw)
    check(lib().pmoe_eca_bn_bwd_apply(C.byref(vdy), C.byref(vx), gate.data_ptr(), gate.stride(0), dmean.data_ptr(), dmean.stride(0),
                                      scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                      b1.data_ptr(), b2.data_ptr(), 1.0 / N, C.byref(vdr), None, stream_ptr()), "eca apply")
    o1, o2 = orthogonality(draw)
    print("   eca + BatchNorm apply: |sum dx| / sum|dx| = %.2e, |sum dx xhat| / sum|dx| = %.2e" % (o1, o2))
    assert o1 < 1e-3 and o2 < 1e-3, (o1, o2)


@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 8, 6, 16), (2, 32, 32, 128)])
def test_stem_tail_backward_through_the_upstream_batchnorm(n, h, w, c):
    """pmoe_bn2_relu_maxpool_bwd_apply: raw -> BN_up (batch statistics) -> ReLU -> [stored as bf16 = x] -> bn1 (batch statistics)
    -> ReLU -> MaxPool2d(3, 2, 1), backward to raw in ONE pass with the upstream sums in closed form, against torch autograd in fp64
    through the same chain (bf16 storage of x as a straight-through rounding). Negative gammas in both BatchNorms."""
    from pmoe_b200 import train
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    g = torch.Generator().manual_seed(13 * n + c)
    raw = (torch.randn(n, h, w, c, generator=g) * 1.3 - 0.2).to(torch.bfloat16).to(dev)
    ug = torch.randn(c, generator=g).to(dev)
    ub = (torch.randn(c, generator=g) * 0.5).to(dev)
    g1 = torch.randn(c, generator=g).to(dev)
    b1 = (torch.randn(c, generator=g) * 0.5).to(dev)
    dp = torch.randn(n, h // 2, w // 2, c, generator=g).to(torch.bfloat16).to(dev)
    N = n * h * w
    rd = raw.double()
    umean = rd.mean(dim=(0, 1, 2))
    urstd = 1.0 / torch.sqrt(rd.var(dim=(0, 1, 2), unbiased=False) + 1e-5)
    uscale = (ug.double() * urstd).float()
    ushift = (ub.double() - umean * ug.double() * urstd).float()
    x = torch.relu(torch.addcmul(ushift, raw.float(), uscale)).to(torch.bfloat16)        # what the forward stores
    xd = x.double()
    mean = xd.mean(dim=(0, 1, 2))
    rstd = 1.0 / torch.sqrt(xd.var(dim=(0, 1, 2), unbiased=False) + 1e-5)
    scale = (g1.double() * rstd).float()
    shift = (b1.double() - mean * g1.double() * rstd).float()
    meanf, rstdf, umeanf, urstdf = mean.float(), rstd.float(), umean.float(), urstd.float()
    p, idx, xm = _fused_forward(x, scale, shift)
    s1 = torch.zeros(c, dtype=torch.float64, device=dev)
    s2 = torch.zeros(c, dtype=torch.float64, device=dev)
    e1 = torch.zeros(c, dtype=torch.float64, device=dev)
    vdp = view4(dp)
    check(lib().pmoe_bn_relu_maxpool_bwd_reduce(C.byref(vdp), xm.data_ptr(), scale.data_ptr(), shift.data_ptr(), meanf.data_ptr(),
                                                rstdf.data_ptr(), s1.data_ptr(), s2.data_ptr(), e1.data_ptr(), stream_ptr()), "reduce")
    # closed-form sums of the upstream BatchNorm (train.bn_relu_maxpool_op)
    ssum, ssq, npos = xd.sum(dim=(0, 1, 2)), (xd * xd).sum(dim=(0, 1, 2)), (x > 0).double().sum(dim=(0, 1, 2))
    g64, r64, m64 = g1.double(), rstdf.double(), meanf.double()
    c1, c2 = s1 / N, s2 / N
    A, Bq, Cq = g64 * r64, -g64 * r64 * r64 * c2, g64 * r64 * (r64 * c2 * m64 - c1)
    n1 = A * e1 + Bq * ssum + Cq * npos
    n2 = A * (s2 / r64 + m64 * s1) + Bq * ssq + Cq * ssum
    usc, ush = uscale.double(), ushift.double()
    sraw = torch.where(usc != 0, (n2 - ush * n1) / torch.where(usc != 0, usc, torch.ones_like(usc)), torch.zeros_like(usc))
    u1 = n1.contiguous()
    u2 = (urstdf.double() * (sraw - umeanf.double() * n1)).contiguous()
    draw = torch.empty_like(raw)
    dgam, dbet, udgam, udbet = [torch.zeros(c, device=dev) for _ in range(4)]
    pg, upg = train.BnParamGrads(), train.BnParamGrads()
    pg.dgamma, pg.dbeta, pg.n, pg.accumulate = dgam.data_ptr(), dbet.data_ptr(), c, 0
    upg.dgamma, upg.dbeta, upg.n, upg.accumulate = udgam.data_ptr(), udbet.data_ptr(), c, 0
    vr, vd = view4(raw), view4(draw)
    check(lib().pmoe_bn2_relu_maxpool_bwd_apply(
        C.byref(vdp), idx.data_ptr(), C.byref(vr), scale.data_ptr(), shift.data_ptr(), meanf.data_ptr(), rstdf.data_ptr(), g1.data_ptr(),
        s1.data_ptr(), s2.data_ptr(), 1.0 / N, C.byref(pg), uscale.data_ptr(), ushift.data_ptr(), umeanf.data_ptr(), urstdf.data_ptr(),
        ug.data_ptr(), u1.data_ptr(), u2.data_ptr(), C.byref(upg), C.byref(vd), stream_ptr()), "bn2 apply")
    # fp64 autograd through the whole chain
    rr = raw.double().requires_grad_(True)
    ugd, ubd = ug.double().requires_grad_(True), ub.double().requires_grad_(True)
    g1d, b1d = g1.double().requires_grad_(True), b1.double().requires_grad_(True)
    mu = rr.mean(dim=(0, 1, 2))
    var = rr.var(dim=(0, 1, 2), unbiased=False)
    xx = torch.relu((rr - mu) / torch.sqrt(var + 1e-5) * ugd + ubd)
    xx = xx + (x.double() - xx).detach()                                   # bf16 storage of x, straight through
    xx.retain_grad()
    mu1 = xx.mean(dim=(0, 1, 2))
    var1 = xx.var(dim=(0, 1, 2), unbiased=False)
    z = torch.relu((xx - mu1) / torch.sqrt(var1 + 1e-5) * g1d + b1d)
    pp = torch.nn.functional.max_pool2d(z.permute(0, 3, 1, 2), 3, 2, 1)
    pp.backward(dp.double().permute(0, 3, 1, 2))
    e = _rel(draw.float(), rr.grad)
    print("\n   two-BatchNorm stem backward %dx%dx%dx%d: d raw %.2e, bn1 %.1e / %.1e, upstream BN %.1e / %.1e" % (
        n, h, w, c, e, _rel(dgam, g1d.grad), _rel(dbet, b1d.grad), _rel(udgam, ugd.grad), _rel(udbet, ubd.grad)))
    assert e < 6e-3                                                          # bf16 storage of d raw (and of x inside the chain)
    assert _rel(dgam, g1d.grad) < 1e-4 and _rel(dbet, b1d.grad) < 1e-4
    assert _rel(udbet, ubd.grad) < 1e-3
    # The upstream d gamma = sum dx*m*xhat_up is recovered from sum dx*x through the forward affine (x = gamma_up*xhat_up + beta_up where
    # x > 0), i.e. through x's bf16 STORAGE: every term carries x's rounding error (<= 2^-9 |x|) divided by gamma_up — the same route
    # conv_op takes for sums handed down by any consumer. Per channel: a few 2^-9 * sqrt(sum (dx*x)^2) / |gamma_up|.
    tol = 6 * 2.0 ** -9 * (xx.grad * x.double()).pow(2).sum(dim=(0, 1, 2)).sqrt() / ug.double().abs().clamp_min(1e-12) \
        + 1e-4 * ugd.grad.abs()
    assert bool(((udgam.double() - ugd.grad).abs() <= tol).all()), ((udgam.double() - ugd.grad).abs() / tol).max().item()
