"""Kernel-level GPU parity of the tensor-core conv variants that the small module tests (32x32 inputs) do not reach:
streamed-weights halo kernel (>= 256 channels at >= 18x10), fused 2x2 max-pool, fp32 NCHW copy, per-image (ECA-gated)
weights, ConvTranspose2d as one multi-view GEMM. Reference: torch fp32 on the same bf16-rounded operands; tolerance 1e-2
relative (bf16 output rounding), pooled / copied outputs must be consistent with the stored output bit for bit."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


def _pack(w, cins, cpads, taps, cop):
    from pmoe_b200 import ops
    return ops.pack_conv_weight(w, cins, cpads, taps, cop)


def _nhwc(x, cpad):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, cpad, dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def _conv_case(B, H, W, cin, cout, gen):
    from pmoe_b200 import ops
    x = torch.randn(B, cin, H, W, generator=gen).to(dev)
    w = (torch.randn(cout, cin, 3, 3, generator=gen) / (cin * 9) ** 0.5).to(dev)
    shift = (torch.randn(cout, generator=gen) * 0.1).to(dev)
    cp, cop, cs = ops.pad_ch(cin), ops.cout_padded(cout), ops.pad_ch(cout)
    ck = ops.choose_ck([cp])
    wp = _pack(w, [cin], [cp], ops.TAPS3, cop)
    segs = ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], [cp], ck)
    ref = torch.relu(F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), padding=1) + shift.view(1, -1, 1, 1))
    return x, w, shift, cp, cop, cs, ck, wp, segs, ref


@pytest.mark.parametrize("cin,cout,H,W", [(256, 256, 32, 24), (256, 128, 36, 20), (512, 512, 20, 16), (128, 256, 24, 24)])
def test_streamed_weight_halo_kernel(cin, cout, H, W):
    from pmoe_b200 import ops
    gen = torch.Generator().manual_seed(cin + cout)
    x, w, shift, cp, cop, cs, ck, wp, segs, ref = _conv_case(2, H, W, cin, cout, gen)
    out = torch.full((2, H, W, cs), 7.0, dtype=torch.bfloat16, device=dev)
    ops.conv_tc([_nhwc(x, cp)], wp, segs, ck, out, None, ops.pad_vec(shift, cop), "relu")
    got = out[..., :cout].permute(0, 3, 1, 2).float()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 1e-2, rel


def test_fused_pool_and_nchw_copy_are_consistent_with_the_stored_output():
    from pmoe_b200 import ops
    gen = torch.Generator().manual_seed(5)
    B, H, W, cin, cout = 2, 48, 32, 64, 64
    x, w, shift, cp, cop, cs, ck, wp, segs, ref = _conv_case(B, H, W, cin, cout, gen)
    out = torch.empty(B, H, W, cs, dtype=torch.bfloat16, device=dev)
    pooled = torch.full((B, H // 2, W // 2, cs), -1.0, dtype=torch.bfloat16, device=dev)
    ops.conv_tc([_nhwc(x, cp)], wp, segs, ck, out, None, ops.pad_vec(shift, cop), "relu", pool2_out=pooled)
    assert ((out[..., :cout].permute(0, 3, 1, 2).float() - ref).norm() / ref.norm()).item() < 1e-2
    want = F.max_pool2d(out.permute(0, 3, 1, 2).float(), 2, 2).permute(0, 2, 3, 1)
    assert torch.equal(pooled.float(), want)  # max of the stored bf16 values: exact
    # 1x1 conv with an fp32 NCHW copy of the first 23 channels
    w1 = (torch.randn(23, 64, 1, 1, generator=gen) / 8).to(dev)
    b1 = torch.randn(23, generator=gen).to(dev)
    cop1 = ops.cout_padded(23)
    wp1 = _pack(w1, [64], [64], [(0, 0)], cop1)
    o1 = torch.empty(B, H, W, ops.pad_ch(23), dtype=torch.bfloat16, device=dev)
    nchw = torch.full((B, 23, H, W), 9.0, dtype=torch.float32, device=dev)
    pool = torch.zeros(B, cop1, dtype=torch.float32, device=dev)
    ops.conv_tc([out], wp1, ops.conv_segments([(0, 0)], [64], 64), 64, o1, None, ops.pad_vec(b1, cop1), None, pool_sum=pool, nchw_out=nchw)
    ref1 = F.conv2d(out[..., :64].permute(0, 3, 1, 2).float(), w1.to(torch.bfloat16).float()) + b1.view(1, -1, 1, 1)
    assert ((nchw - ref1).norm() / ref1.norm()).item() < 1e-3       # fp32 copy: unrounded accumulator + bias
    assert ((o1[..., :23].permute(0, 3, 1, 2).float() - ref1).norm() / ref1.norm()).item() < 1e-2
    psum = o1.float().sum(dim=(1, 2))
    assert ((pool[:, :23] - psum[:, :23]).norm() / psum[:, :23].norm()).item() < 1e-4  # per-image sums of the STORED values


def test_per_image_gated_weights_equal_scaling_the_input():
    from pmoe_b200 import ops
    gen = torch.Generator().manual_seed(9)
    B, H, W, cin, cout = 5, 40, 24, 128, 64   # several images per CTA range: weights are reloaded at image boundaries
    x, w, shift, cp, cop, cs, ck, wp, segs, _ = _conv_case(B, H, W, cin, cout, gen)
    gate = torch.rand(B, cp, generator=gen).to(dev)
    wimg = ops.gate_weights(wp, gate, cp)
    out = torch.empty(B, H, W, cs, dtype=torch.bfloat16, device=dev)
    ops.conv_tc([_nhwc(x, cp)], wimg, segs, ck, out, None, ops.pad_vec(shift, cop), "relu")
    xs = x.to(torch.bfloat16).float() * gate[:, :cin].view(B, cin, 1, 1)
    ref = torch.relu(F.conv2d(xs, w.to(torch.bfloat16).float(), padding=1) + shift.view(1, -1, 1, 1))
    assert ((out[..., :cout].permute(0, 3, 1, 2).float() - ref).norm() / ref.norm()).item() < 1e-2


def test_conv_transpose_as_one_gemm_matches_torch():
    from pmoe_b200 import config, infer
    from pmoe_b200.nhwc import Act
    gen = torch.Generator().manual_seed(3)
    up = torch.nn.ConvTranspose2d(128, 64, kernel_size=2, stride=2).to(dev)
    x = torch.randn(2, 128, 20, 12, generator=gen).to(dev)
    with torch.no_grad(), config.use_precision("bf16"):
        y = infer.conv_transpose_eval(up, Act(_nhwc(x, 128), 128))
        ref = F.conv_transpose2d(x.to(torch.bfloat16).float(), up.weight.to(torch.bfloat16).float(), up.bias, stride=2)
    got = y.t[..., :64].permute(0, 3, 1, 2).float()
    assert ((got - ref).norm() / ref.norm()).item() < 1e-2


@pytest.mark.parametrize("cin,cout,hw", [(64, 128, 32), (128, 256, 16), (256, 512, 8), (64, 64, 28)])
def test_stride2_dgrad_fused_matches_per_parity_launches(cin, cout, hw):
    """Data gradient of a 3x3/stride-2 conv (ResNet layer2-4 entry blocks) as ONE GEMM over the dy grid (K = 4 shifts x Cout,
    N = 4 parities x Cin, stored through two row-parity views) and as four per-parity-view launches, each followed by the
    1x1/stride-2 downsample branch accumulating into the same gradient buffer: both against torch autograd on the same
    bf16-rounded operands (north_star's bf16 tolerance, 1e-2; measured ~2e-3). BatchNorm is left out on purpose: at test sizes
    its backward amplifies bf16 rounding differences between two correct accumulation orders to tens of percent."""
    import torch.nn as nn
    from pmoe_b200 import config, train

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.c0 = nn.Conv2d(16, cin, 1, bias=False)
            self.c1 = nn.Conv2d(cin, cout, 3, 2, 1, bias=False)
            self.down = nn.Conv2d(cin, cout, 1, 2, bias=False)

        def forward(self, x):
            def body(tape, a):
                y, _ = train.conv_op(tape, [a], self.c0.weight, None, None, None, ksize=1, tag="c0")
                s1, d1 = train.stride2_sources(y, 1)
                idt, _ = train.conv_op(tape, s1, self.down.weight, None, None, None, segdefs=d1,
                                       out_hw=(s1[0].t.shape[1], s1[0].t.shape[2]), tag="down")
                s3, d3 = train.stride2_sources(y, 3)
                z, _ = train.conv_op(tape, s3, self.c1.weight, None, None, None, segdefs=d3, residual=idt,
                                     out_hw=(s3[0].t.shape[1], s3[0].t.shape[2]), tag="c1", stride2_of=y)
                return z
            return train.nhwc_module_forward(self, x, body)

    torch.manual_seed(cin + hw)
    m = Block().cuda()
    x = torch.randn(3, 16, hw, hw, device="cuda")
    cot = torch.randn(3, cout, hw // 2, hw // 2, device="cuda")
    w0 = m.c0.weight.detach().bfloat16().float().requires_grad_(True)
    w1 = m.c1.weight.detach().bfloat16().float().requires_grad_(True)
    wd = m.down.weight.detach().bfloat16().float().requires_grad_(True)
    y = F.conv2d(x.bfloat16().float(), w0)
    z = F.conv2d(y, w1, None, 2, 1) + F.conv2d(y, wd, None, 2)
    (z * cot).sum().backward()

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    got = {}
    old = train.FUSE_STRIDE2_DGRAD
    try:
        for fuse in (True, False):
            train.FUSE_STRIDE2_DGRAD = fuse
            m.zero_grad()
            with config.use_precision("bf16"):
                out = m(x)
                (out * cot).sum().backward()
            assert rel(out, z.detach()) < 1e-2
            # d(c0.weight) is a function of the data gradient under test; the other two pin the rest of the block
            for name, ref in (("c0", w0.grad), ("c1", w1.grad), ("down", wd.grad)):
                e = rel(getattr(m, name).weight.grad, ref)
                assert e < 1e-2, (fuse, name, e)
            got[fuse] = m.c0.weight.grad.clone()
    finally:
        train.FUSE_STRIDE2_DGRAD = old
    assert rel(got[True], got[False]) < 5e-3
