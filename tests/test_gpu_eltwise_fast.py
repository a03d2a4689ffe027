"""Dense-bf16 fast paths of the memory-bound kernels against torch on the same bf16 operands (through the C-ABI):
channel_sums / channel_stats (per-image and per-channel reductions), affine_act_stats (BN-apply + ReLU fused with the
statistics of the stored output), eca_bwd_apply. Each is also run on a strided view, which takes the generic kernel, and
the two must agree. Sums are fp32 block partials: 1e-5 relative on well-conditioned (non-cancelling) inputs."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 7, 9, 48), (5, 4, 4, 512), (1, 33, 17, 16)])
def test_channel_sums_and_stats_fast_vs_torch(n, h, w, c):
    from pmoe_b200 import nhwc
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    g = torch.Generator().manual_seed(n * 100 + c)
    x = (torch.rand(n, h, w, c, generator=g) + 0.25).to(dev).to(torch.bfloat16)
    wide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device=dev)
    wide[..., :c] = x
    xs = wide[..., :c]                       # strided view -> generic kernel
    ref = x.float().sum(dim=(1, 2))
    assert _rel(nhwc.channel_sums(x), ref) < 1e-5
    assert _rel(nhwc.channel_sums(xs), ref) < 1e-5
    for t in (x, xs):
        s1 = torch.zeros(c, dtype=torch.float64, device=dev)
        s2 = torch.zeros(c, dtype=torch.float64, device=dev)
        v = view4(t)
        check(lib().pmoe_channel_stats(C.byref(v), nhwc.dtype_code(t), s1.data_ptr(), s2.data_ptr(), stream_ptr()), "channel_stats")
        assert _rel(s1, x.double().sum(dim=(0, 1, 2))) < 1e-5
        assert _rel(s2, (x.double() ** 2).sum(dim=(0, 1, 2))) < 1e-5


@pytest.mark.parametrize("act", [None, "relu"])
@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 7, 9, 48), (4, 14, 14, 256)])
def test_affine_act_stats_matches_separate_kernels(act, n, h, w, c):
    from pmoe_b200 import nhwc
    g = torch.Generator().manual_seed(7 * n + c)
    x = torch.randn(n, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    scale = (torch.rand(c, generator=g) + 0.5).to(dev)
    shift = (torch.randn(c, generator=g) * 0.3).to(dev)
    ref = nhwc.affine_act(x, scale, shift, act)
    ref_f = x.float() * scale + shift
    if act == "relu":
        ref_f = ref_f.clamp_min(0)
    assert torch.equal(ref, ref_f.to(torch.bfloat16)) or _rel(ref.float(), ref_f) < 4e-3
    for want_pool, want_stats in ((True, False), (False, True), (True, True)):
        out = torch.empty_like(x)
        pool = torch.zeros(n, c + 16, dtype=torch.float32, device=dev) if want_pool else None
        stats = (torch.zeros(c, dtype=torch.float64, device=dev), torch.zeros(c, dtype=torch.float64, device=dev)) if want_stats else None
        assert nhwc.affine_act_stats(x, scale, shift, act, out, pool, 0, stats)
        assert torch.equal(out, ref)                      # same stored bf16 values as the plain kernel
        if want_pool:                                      # statistics of the STORED values
            assert _rel(pool[:, :c], ref.float().sum(dim=(1, 2))) < 1e-5 and pool[:, c:].abs().max().item() == 0
        if want_stats:
            ref_sum = ref.double().sum(dim=(0, 1, 2))   # may cancel without the ReLU: bound relative to the sum of magnitudes
            assert (stats[0] - ref_sum).abs().max().item() <= 1e-5 * ref.double().abs().sum(dim=(0, 1, 2)).max().item()
            assert _rel(stats[1], (ref.double() ** 2).sum(dim=(0, 1, 2))) < 1e-5
    # a strided output does not qualify: nothing is launched and the caller falls back
    wide = torch.empty(n, h, w, 2 * c, dtype=torch.bfloat16, device=dev)
    assert nhwc.affine_act_stats(x, scale, shift, act, wide[..., :c], torch.zeros(n, c, device=dev), 0, None) is False


@pytest.mark.parametrize("accumulate", [False, True])
@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 7, 9, 48), (2, 14, 14, 16)])
def test_eca_bwd_apply_fast_vs_generic(accumulate, n, h, w, c):
    from pmoe_b200 import nhwc
    from pmoe_b200._lib import lib, check, view4, stream_ptr
    g = torch.Generator().manual_seed(3 * n + c)
    dout = torch.randn(n, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    gate = torch.rand(n, c, generator=g).to(dev)
    dmean = (torch.randn(n, c, generator=g) * 0.1).to(dev)
    base = torch.randn(n, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    ref = dout.float() * gate[:, None, None, :] + dmean[:, None, None, :] + (base.float() if accumulate else 0)
    outs = []
    for strided in (False, True):
        if strided:
            wide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device=dev)
            dx = wide[..., :c]
            dx.copy_(base)
        else:
            dx = base.clone()
        vd, vx = view4(dout), view4(dx)
        check(lib().pmoe_eca_bwd_apply(C.byref(vd), nhwc.dtype_code(dout), gate.data_ptr(), gate.stride(0), dmean.data_ptr(),
                                       dmean.stride(0), C.byref(vx), int(accumulate), stream_ptr()), "eca_bwd_apply")
        outs.append(dx.float().clone())
        assert ((dx.float() - ref).abs() <= ref.abs().clamp_min(1e-2) * 2.0 ** -7).all()
    assert _rel(outs[0], outs[1]) < 4e-3


@pytest.mark.parametrize("n,h,w,c", [(3, 20, 28, 64), (2, 7, 9, 48), (2, 14, 14, 16)])
def test_scale_channels_fast_vs_generic(n, h, w, c):
    from pmoe_b200 import nhwc
    g = torch.Generator().manual_seed(11 * n + c)
    x = torch.randn(n, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    gate = torch.rand(n, c, generator=g).to(dev)
    ref = (x.float() * gate[:, None, None, :]).to(torch.bfloat16)
    assert torch.equal(nhwc.scale_channels(x, gate), ref)
    wide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device=dev)
    wide[..., :c] = x
    assert torch.equal(nhwc.scale_channels(wide[..., :c], gate), ref)   # strided source: generic kernel


@pytest.mark.parametrize("c", [3, 12, 16, 23, 36])
def test_nchw_to_nhwc_narrow_and_wide(c):
    from pmoe_b200 import nhwc
    g = torch.Generator().manual_seed(c)
    x = torch.rand(3, c, 18, 22, generator=g).to(dev)
    for dtype in (torch.bfloat16, torch.float32):
        a = nhwc.from_nchw(x, dtype=dtype)
        cp = a.t.shape[3]
        assert cp % 16 == 0 and a.c == c
        assert torch.equal(a.t[..., :c], x.permute(0, 2, 3, 1).to(dtype))
        assert a.t[..., c:].abs().sum().item() == 0
    xs = torch.rand(3, 2 * c, 18, 22, generator=g).to(dev)[:, ::2]      # strided source planes
    a = nhwc.from_nchw(xs, dtype=torch.bfloat16)
    assert torch.equal(a.t[..., :c], xs.permute(0, 2, 3, 1).to(torch.bfloat16))


def test_bn_chain_sums_fused_into_apply_matches_reduce_pass():
    """torchvision's bn1 follows the stem block's conv+BN+ReLU directly: bn1's backward-apply kernel also reduces the two
    backward sums of the UPSTREAM BatchNorm (sum dx*[x>0], sum dx*x -> sum dy*m and sum dy*m*raw through the forward affine),
    which then needs no reduce pass. Every gradient of the stem must agree with the separate-reduce form to bf16 noise, and
    both with the fp32 CUDA-core path."""
    import torch.nn as nn
    from pmoe_b200 import config, train
    from pmoe_b200.model.blocks.basics import EfficientConvBlock

    class Stem(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = EfficientConvBlock(12, 64)
            self.bn1 = nn.BatchNorm2d(64)

        def forward(self, x):
            def body(tape, a):
                y = train.eca_conv_block(tape, self.conv1, a, tag="stem.conv1", want_out_stats=True)
                y = train.bn_act_op(tape, self.bn1, y, "relu", tag="stem.bn1")
                return train.maxpool_op(tape, y, 3, 2, 1)
            return train.nhwc_module_forward(self, x, body)

    torch.manual_seed(5)
    ref_mod = Stem()
    with torch.no_grad():
        for m_ in ref_mod.modules():
            if isinstance(m_, nn.BatchNorm2d):
                m_.weight.uniform_(0.5, 1.5)
                m_.bias.normal_(0, 0.2)
    sd = ref_mod.state_dict()
    x = torch.rand(8, 12, 64, 64, device=dev)
    cot = torch.randn(8, 64, 32, 32, device=dev) * 1e-2

    def run(prec, fuse):
        old = train.FUSE_BN_CHAIN_SUMS
        train.FUSE_BN_CHAIN_SUMS = fuse
        try:
            with config.use_precision(prec):
                m = Stem()
                m.load_state_dict(sd)
                m = m.cuda().train()
                out = m(x)
                (out * cot).sum().backward()
                return out.detach().float(), {n: p.grad.detach().double().clone() for n, p in m.named_parameters()}
        finally:
            train.FUSE_BN_CHAIN_SUMS = old

    (of, gf), (os_, gs), (o32, g32) = run("bf16", True), run("bf16", False), run("fp32", False)
    # the forward is untouched by the switch; two passes may still differ by single bf16 ulps on a few elements (the per-image
    # ECA pool sums are fp32 atomics, whose order is not fixed from launch to launch)
    assert ((of - os_).norm() / os_.norm()).item() < 1e-3 and (of != os_).float().mean().item() < 1e-2
    names = [n for n in gf if "eca" not in n]                     # ECA kernel gradients: cancelling sums, not a yardstick
    rows = []
    for n in names:
        d = (gf[n] - gs[n]).norm().item() / max(gs[n].norm().item(), 1e-30)
        ef = (gf[n] - g32[n]).norm().item() / max(g32[n].norm().item(), 1e-30)
        es = (gs[n] - g32[n]).norm().item() / max(g32[n].norm().item(), 1e-30)
        rows.append((n, d, ef, es))
    print("\n" + "\n".join("%-34s fused-vs-split %.2e | fused-vs-fp32 %.2e | split-vs-fp32 %.2e" % r for r in rows))
    for n, d, ef, es in rows:
        # the fused sums see the upstream raw output through its bf16-rounded ReLU output (one more rounding than the separate
        # pass): the two forms may differ by bf16 noise, and the fused one must be as close to the fp32 path as the split one
        assert d < 3e-2, (n, d)
        assert ef < max(3e-2, 1.5 * es), (n, ef, es)
