"""Eval-mode (inference) execution of the reference modules on the B200 kernels.

BatchNorm is folded into the conv epilogue (scale/shift), ReLU / bias / residual add ride the same
epilogue, skip-concats and the PU-Net mask deque are channel-slice views (no copies), and
ConvTranspose2d k2s2 writes straight into pixel-shuffle views. Packed bf16 weights and folded
affines are derived caches keyed on the parameters' versions — never part of the state_dict.
"""
import torch

from . import config, nhwc, ops
from .nhwc import Act
from .ops import TAPS3, pad_ch


# ------------------------------------------------------------------------------- derived-tensor cache
def cached(owner, name, deps, build):
    cache = owner.__dict__.setdefault("_pmoe_cache", {})
    key = tuple((t.data_ptr(), t._version, str(t.device)) for t in deps)
    hit = cache.get(name)
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        val = build()
    cache[name] = (key, val)
    return val


def folded_bn(bn, cout_pad):
    """Eval-mode BatchNorm as a per-channel affine (basics.py:52,55 in .eval())."""
    def build():
        scale = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
        shift = bn.bias.float() - bn.running_mean.float() * scale
        return ops.pad_vec(scale, cout_pad, 0.0), ops.pad_vec(shift, cout_pad, 0.0)
    return cached(bn, "fold%d" % cout_pad, [bn.weight, bn.bias, bn.running_mean, bn.running_var], build)


def packed_conv(conv, groups, group_pads, taps, cout_pad, name="w", bn=None):
    """conv.weight (Cout, sum(groups), R, S) -> [cout_pad, K] bf16 in (tap, group, padded channel) order. With `bn`,
    the eval-mode BatchNorm scale gamma/sqrt(var+eps) is folded into the rows (the epilogue then adds the shift only)."""
    key = "%s_%s_%s_%d_%s_%s" % (name, "-".join(map(str, groups)), "-".join(map(str, group_pads)), cout_pad, config.precision(),
                                 "bn" if bn is not None else "")
    deps = [conv.weight] if bn is None else [conv.weight, bn.weight, bn.running_var]

    def build():
        w = conv.weight.detach().float()
        if bn is not None:
            w = w * (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)).view(-1, 1, 1, 1)
        return ops.pack_conv_weight(w, groups, group_pads, taps, cout_pad, config.act_dtype())
    return cached(conv, key, deps, build)


def padded_bias(mod, cout_pad):
    return cached(mod, "bias%d" % cout_pad, [mod.bias], lambda: ops.pad_vec(mod.bias.detach().float(), cout_pad, 0.0))


# ------------------------------------------------------------------------------- layers
def _tc():
    return config.precision() == "bf16" and not config.FORCE_SIMT


def conv_eval(srcs, conv, bn=None, act=None, out=None, residual=None, pool=None, pool_stride=0, groups=None, ksize=3,
              pool2=False, gate=None):
    """srcs: list of Act (virtual concat along channels). groups: optional [(logical, padded)] layout of
    the channel axis when one physical source carries several padded groups (the PU-Net mask ring)."""
    cout = conv.weight.shape[0]
    cop = ops.cout_padded(cout)
    if groups is None:
        glog = [a.c for a in srcs]
        gpad = [a.cpad for a in srcs]
    else:
        glog = [g[0] for g in groups]
        gpad = [g[1] for g in groups]
    phys = [a.cpad for a in srcs]
    assert sum(gpad) == sum(phys), (gpad, phys)
    assert sum(glog) == conv.weight.shape[1], (glog, conv.weight.shape)
    ck = ops.choose_ck(phys)
    taps = TAPS3 if ksize == 3 else [(0, 0)]
    pad = ksize // 2
    wp = packed_conv(conv, glog, gpad, taps, cop, bn=bn)
    if gate is not None:  # per-image weights carrying an ECA gate over the (single) source's physical channels
        assert len(srcs) == 1
        wp = ops.gate_weights(wp, gate, phys[0])
    segs = ops.conv_segments([(r - pad, s - pad) for (r, s) in taps], phys, ck)
    scale = None  # the BatchNorm scale lives in the packed weights
    if bn is not None:
        shift = folded_bn(bn, cop)[1]
    else:
        shift = padded_bias(conv, cop) if conv.bias is not None else None
    n, h, w, _ = srcs[0].t.shape
    if out is None:
        out = torch.empty(n, h, w, pad_ch(cout), dtype=config.act_dtype(), device=srcs[0].t.device)
    pooled = None
    kw = {}
    if pool2 and _tc():  # MaxPool2d(2,2) of this output rides the conv epilogue (unet.py:29)
        pooled = torch.empty(n, h // 2, w // 2, out.shape[3], dtype=out.dtype, device=out.device)
        kw["pool2_out"] = pooled
    ops.conv([a.t for a in srcs], wp, segs, ck, out, scale=scale, shift=shift, act=act,
             residual=None if residual is None else residual.t, pool_sum=pool, pool_stride=pool_stride,
             flops=2.0 * n * h * w * cout * sum(glog) * len(taps), tag="%dx%d %d->%d k%d" % (h, w, sum(glog), cout, ksize), **kw)
    if pool2:
        return Act(out, cout), (Act(pooled, cout) if pooled is not None else nhwc.maxpool(Act(out, cout), 2, 2, 0))
    return Act(out, cout)


def conv3_block_eval(seq, srcs, pool=None, pool2=False):
    """conv3 (basics.py:48-59) in eval mode: two fused conv+BN+ReLU launches; pool2 also returns MaxPool2d(2,2)(out)."""
    y = conv_eval(srcs, seq[0], seq[1], "relu")
    return conv_eval([y], seq[3], seq[4], "relu", pool=pool, pool2=pool2)


def conv_transpose_eval(up, x):
    """nn.ConvTranspose2d(k=2, s=2) (unet.py:35-45) as ONE GEMM with N = 4*Cout: column block q = 2a+b holds the
    weights of output parity (a, b) and is stored into the pixel-shuffle view out[:, a::2, b::2, :]."""
    cin, cout = up.weight.shape[0], up.weight.shape[1]
    n, h, w, cp = x.t.shape
    cs = pad_ch(cout)
    out = torch.empty(n, 2 * h, 2 * w, cs, dtype=config.act_dtype(), device=x.t.device)
    ck = ops.choose_ck([cp])
    segs = ops.conv_segments([(0, 0)], [cp], ck)
    views = [out[:, a::2, b::2, :] for a in range(2) for b in range(2)]
    if cs % 64 == 0 and _tc():
        # column blocks (2a, 2a+1) = the two horizontal parities of output row 2h+a are adjacent pixels: seen as ONE
        # (n, h, w, 2*cs) tensor per row parity a, whose pixel rows are 2*cs contiguous channels -> two plain views
        wp, shift = cached(up, "wq4_%d_%d_%s" % (cp, cs, config.precision()), [up.weight, up.bias], lambda: ops.pack_convT_weight(up, cp, cs))
        rows = out.view(n, h, 2, w, 2 * cs)
        ops.conv([x.t], wp, segs, ck, rows[:, :, 0], shift=shift, out_extra=[rows[:, :, 1]], out_cols=2 * cs,
                 flops=2.0 * n * h * w * cin * cout * 4, tag="convT %dx%d %d->%d" % (h, w, cin, cout))
        return Act(out, cout)
    cop = ops.cout_padded(cout)
    shift = padded_bias(up, cop)
    for q, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        wp = cached(up, "wq%d%d_%d_%d_%s" % (a, b, cp, cop, config.precision()), [up.weight],
                    lambda: ops.pack_conv_weight(up.weight.detach().float()[:, :, a, b].t().reshape(cout, cin, 1, 1),
                                                 [cin], [cp], [(0, 0)], cop, config.act_dtype()))
        ops.conv([x.t], wp, segs, ck, views[q], shift=shift,
                 flops=2.0 * n * h * w * cin * cout, tag="convT %dx%d %d->%d" % (h, w, cin, cout))
    return Act(out, cout)


def unet_eval(net, x, out=None, out_pool=None, pool_stride=0, want_inter=False, nchw_out=None):
    """UNet.forward (unet.py:50-95) in eval mode. x: Act (N,H,W,16). Logits go to `out` (NHWC bf16
    view with >= 32 channels) if given. Returns (Act logits, inter or None)."""
    n, h, w, _ = x.t.shape
    if h % 16 or w % 16:
        raise RuntimeError("pmoe_b200 UNet needs H and W divisible by 16 (got %dx%d)" % (h, w))
    x1, p1 = conv3_block_eval(net.dwn_1, [x], pool2=True)
    x2, p2 = conv3_block_eval(net.dwn_2, [p1], pool2=True)
    x3, p3 = conv3_block_eval(net.dwn_3, [p2], pool2=True)
    x4, p4 = conv3_block_eval(net.dwn_4, [p3], pool2=True)
    inter_sum = None
    if want_inter:
        inter_sum = torch.zeros(n, 512, dtype=torch.float32, device=x.t.device)
    x5 = conv3_block_eval(net.dwn_5, [p4], pool=inter_sum)
    y = x5
    for up, fwd, skip in ((net.up_1, net.up_forw_1, x4), (net.up_2, net.up_forw_2, x3), (net.up_3, net.up_forw_3, x2),
                          (net.up_4, net.up_forw_4, x1)):
        u = conv_transpose_eval(up, y)
        y = conv3_block_eval(fwd, [skip, u])  # cat([skip, up]) (unet.py:73) as two K segments
    ncls = net.out.weight.shape[0]
    cop = ops.cout_padded(ncls)
    wp = packed_conv(net.out, [64], [64], [(0, 0)], cop)
    if out is None:
        out = torch.empty(n, h, w, pad_ch(ncls), dtype=config.act_dtype(), device=x.t.device)
    kw = {"nchw_out": nchw_out} if (nchw_out is not None and _tc()) else {}
    ops.conv([y.t], wp, ops.conv_segments([(0, 0)], [64], 64), 64, out, shift=padded_bias(net.out, cop), pool_sum=out_pool,
             pool_stride=pool_stride, flops=2.0 * n * h * w * 64 * ncls, tag="out1x1 %dx%d" % (h, w), **kw)
    if nchw_out is not None and not kw:
        nhwc.to_nchw(out, ncls, out=nchw_out)
    inter = None
    if want_inter:
        inter = inter_sum / float(x5.t.shape[1] * x5.t.shape[2])
    return Act(out, ncls), inter


def eca_conv_block_eval(blk, x_t, groups, pool_in, hw):
    """EfficientConvBlock (basics.py:80-135) in eval mode.
    x_t: NHWC bf16 tensor/view; groups = (n_groups, logical, slot) channel layout; pool_in: fp32 per-image
    channel sums of x_t (N, n_groups*slot) — produced for free by the epilogue that wrote x_t.
    On the tensor-core path the two ECA gates never touch the activations: conv(x * gate) = conv with per-image weights
    W[co][k] * gate[n][ci(k)], so only the (small) packed weights are rewritten per image."""
    ng, gl, gs = groups
    gate1 = nhwc.eca_gate(pool_in, hw, blk.layer1.eca1.conv.weight, ng, gl, gs)
    n = x_t.shape[0]
    pool64 = torch.zeros(n, 64, dtype=torch.float32, device=x_t.device)
    fold = _tc() and x_t.shape[3] % 64 == 0 and x_t.shape[1] >= 18 and x_t.shape[2] >= 10
    if fold:
        c1 = conv_eval([Act(x_t, ng * gl)], blk.layer1.conv1[0], blk.layer1.conv1[1], "relu", pool=pool64, groups=[(gl, gs)] * ng,
                       gate=gate1)
        gate2 = nhwc.eca_gate(pool64, hw, blk.layer2.eca2.conv.weight, 1, 64, 64)
        return conv_eval([c1], blk.layer2.conv2[0], blk.layer2.conv2[1], "relu", gate=gate2)
    xs = nhwc.scale_channels(x_t, gate1)
    c1 = conv_eval([Act(xs, ng * gl)], blk.layer1.conv1[0], blk.layer1.conv1[1], "relu", pool=pool64,
                   groups=[(gl, gs)] * ng)
    gate2 = nhwc.eca_gate(pool64, hw, blk.layer2.eca2.conv.weight, 1, 64, 64)
    c1s = nhwc.scale_channels(c1.t, gate2)
    return conv_eval([Act(c1s, 64)], blk.layer2.conv2[0], blk.layer2.conv2[1], "relu")


def pinned_output_like(B, Fu, ncls, H, W):
    """Pinned host buffer for `PredictiveUnet.forward(..., host_out=)`: logically (B, F, classes, H, W) like the module's
    output, stored frame-major so that one future frame of the whole batch is one contiguous block (one plain async copy)."""
    return torch.empty(Fu, B, ncls, H, W, dtype=torch.float32).pin_memory().permute(1, 0, 2, 3, 4)


def punet_eval(net, images, host_out=None, copy_stream=None):
    """PredictiveUnet.forward (punet.py:75-120) in eval mode. images: fp32 (B,T,C,H,W) on the GPU.
    The deque of the last `past_frames` masks is a sliding 4-slot window over one ring buffer
    (B,H,W,(T+F)*32): every U-Net writes its logits into its slot, and the entry block reads the window
    as a single 128-channel source."""
    B, T, Cin, H, W = images.shape
    P, Fu = net.n_past_frames, net.n_future_frames
    if T != P:
        raise AssertionError("Number of images should match number of past frames")
    ncls = net.unet.out.weight.shape[0]
    slot = pad_ch(ncls)
    dev = images.device
    nslots = P + max(Fu, 0)
    ring = torch.empty(B, H, W, nslots * slot, dtype=config.act_dtype(), device=dev)
    pools = torch.zeros(B, nslots * slot, dtype=torch.float32, device=dev)
    inter = None
    for t in range(P):
        x = nhwc.from_nchw(images[:, t], dtype=config.act_dtype())
        _, it = unet_eval(net.unet, x, out=ring[..., t * slot:(t + 1) * slot], out_pool=pools[:, t * slot:],
                          pool_stride=nslots * slot, want_inter=(net.unet_inter_repr and Fu == 0 and t == P - 1))
        inter = it
    if Fu == 0:
        if net.unet_inter_repr:
            return inter
        return nhwc.to_nchw(ring[..., (P - 1) * slot:P * slot], ncls)
    # frame-major storage, returned as the (B, F, classes, H, W) view the reference's torch.stack(dim=1) has: each frame of
    # the batch is one contiguous block, which is what lets it leave for the host while the next U-Net pass runs
    out = None if net.inter_repr else torch.empty(Fu, B, ncls, H, W, dtype=torch.float32, device=dev).permute(1, 0, 2, 3, 4)
    if host_out is not None:
        if out is None or not host_out.is_pinned() or tuple(host_out.shape) != tuple(out.shape) or not host_out[:, 0].is_contiguous():
            raise RuntimeError("pmoe_b200 punet_eval: host_out must come from infer.pinned_output_like(B, F, classes, H, W)")
        copy_stream = copy_stream if copy_stream is not None else torch.cuda.Stream(device=dev)
    for f in range(Fu):
        window = ring[..., f * slot:(f + P) * slot]
        m = eca_conv_block_eval(net.entry_block, window, (P, ncls, slot), pools[:, f * slot:(f + P) * slot], H * W)
        # the fp32 NCHW logits the module returns are written by the same epilogue that fills the ring slot
        _, inter = unet_eval(net.pred_unet, m, out=ring[..., (P + f) * slot:(P + f + 1) * slot],
                             out_pool=pools[:, (P + f) * slot:], pool_stride=nslots * slot, want_inter=net.inter_repr,
                             nchw_out=None if out is None else out[:, f])
        if host_out is not None:
            ready = torch.cuda.Event()
            ready.record()
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                host_out[:, f].copy_(out[:, f], non_blocking=True)
    if net.inter_repr:
        return inter
    return out
