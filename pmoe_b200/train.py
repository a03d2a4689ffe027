"""General (training-capable) execution of the reference modules on the B200 kernels.

A forward pass records a static tape of backward closures; `TapeFunction` exposes the whole model
call as ONE torch.autograd.Function so the reference trainers' `loss.backward()`, `clip_grad_norm_`
and optimizers work unchanged on the real nn.Parameters. Inside the tape everything is our own
kernels: gradient accumulation at fan-out points is fused into the producing kernel (residual /
accumulate arguments), BatchNorm follows each layer's own `.training` flag (so a "frozen" U-Net that
the trainer flipped back to train mode keeps updating its running statistics, exactly like the
reference — punet.py:55 vs train_1.py:123), and frozen up-stream sub-networks do no backward work.
"""
import ctypes as C

import torch

from . import _lib, config, nhwc, ops, profiler
from ._lib import ACT, check, lib, stream_ptr, view4
from .nhwc import Act, dtype_code
from .ops import TAPS3, pad_ch


# ------------------------------------------------------------------------------------------------ tape
class Tape:
    def __init__(self, dtype, save):
        self.dtype = dtype
        self.save = save            # False: forward only (no_grad), nothing is kept for backward
        self.ops = []
        self.grads = {}             # id(Act) -> NHWC gradient tensor
        self.pgrads = {}            # id(param) -> fp32 gradient in the parameter's own layout
        self.params = {}            # id(param) -> param
        self.alive = []             # keeps Acts (and therefore their ids) alive until backward

    def record(self, fn):
        if self.save:
            self.ops.append(fn)

    def track(self, act):
        if self.save:
            self.alive.append(act)
        return act

    def grad_of(self, act):
        return self.grads.pop(id(act), None)

    def add_pgrad(self, p, g):
        if not p.requires_grad:
            return
        k = id(p)
        self.params[k] = p
        g = g.reshape(p.shape).to(torch.float32)
        if k in self.pgrads:
            self.pgrads[k] = self.pgrads[k] + g
        else:
            self.pgrads[k] = g

    def backward(self):
        for fn in reversed(self.ops):
            fn()
        self.ops = []
        self.alive = []
        self.grads = {}


def _new_act(tape, t, c, rg):
    a = Act(t, c)
    a.rg = bool(rg) and tape.save
    return tape.track(a)


def _rg(a):
    return getattr(a, "rg", False)


def _any_rg(params):
    return any(p is not None and p.requires_grad for p in params)


# ------------------------------------------------------------------------------------------------ weight packing
def _layout_of(srcs, layouts):
    """Per physical source: list of (logical, padded) channel groups."""
    out = []
    for i, a in enumerate(srcs):
        out.append([(a.c, a.cpad)] if layouts is None or layouts[i] is None else list(layouts[i]))
    return out


def _pack_fwd(w, layouts, taps, cop, dtype):
    glog = [g[0] for lay in layouts for g in lay]
    gpad = [g[1] for lay in layouts for g in lay]
    return ops.pack_conv_weight(w.detach().float(), glog, gpad, taps, cop, dtype)


def _unpack_wgrad(dwp, w_shape, layouts, ntaps):
    """[cop, ktot] packed gradient -> (cout, cin, R, S)."""
    cout, cin, R, S = w_shape
    gpad_total = sum(g[1] for lay in layouts for g in lay)
    d = dwp[:cout].reshape(cout, ntaps, gpad_total)
    parts, off = [], 0
    for lay in layouts:
        for (gl, gp) in lay:
            parts.append(d[:, :, off:off + gl])
            off += gp
    d = torch.cat(parts, dim=2)  # (cout, ntaps, cin)
    return d.permute(0, 2, 1).reshape(cout, cin, R, S)


def _pack_dgrad(w, lay, cin_begin, taps, co_pad, dtype):
    """Data-gradient weight for ONE physical source whose channels are laid out as `lay`:
    rows = physical input channels, K = (tap, padded cout). Wd[p, t*co_pad+co] = w[co, ci(p), r_t, s_t]."""
    cout = w.shape[0]
    wf = w.detach().float()
    rows = []
    ci = cin_begin
    for (gl, gp) in lay:
        blk = wf[:, ci:ci + gl]                                   # (cout, gl, R, S)
        blk = torch.stack([blk[:, :, r, s] for (r, s) in taps], 0)  # (ntaps, cout, gl)
        blk = blk.permute(2, 0, 1)                                # (gl, ntaps, cout)
        if co_pad > cout:
            blk = torch.nn.functional.pad(blk, (0, co_pad - cout))
        blk = blk.reshape(gl, len(taps) * co_pad)
        if gp > gl:
            blk = torch.nn.functional.pad(blk, (0, 0, 0, gp - gl))
        rows.append(blk)
        ci += gl
    wd = torch.cat(rows, 0)
    rows_pad = ops.cout_padded(wd.shape[0])
    if rows_pad > wd.shape[0]:
        wd = torch.nn.functional.pad(wd, (0, 0, 0, rows_pad - wd.shape[0]))
    return wd.to(dtype).contiguous()


# ------------------------------------------------------------------------------------------------ raw launch helpers
def _bn_bwd_reduce(dz, z, x, act, mean, rstd, cpad):
    s1 = torch.zeros(cpad, dtype=torch.float32, device=dz.device)
    s2 = torch.zeros(cpad, dtype=torch.float32, device=dz.device) if x is not None else None
    vdz = view4(dz)
    vz = view4(z) if z is not None else _lib.null_view()
    vx = view4(x) if x is not None else _lib.null_view()
    check(profiler.launch("bn_bwd_reduce", lambda: lib().pmoe_bn_bwd_reduce(
        C.byref(vdz), C.byref(vz), C.byref(vx), dtype_code(dz), ACT[act], _lib.ptr(mean), _lib.ptr(rstd), s1.data_ptr(),
        _lib.ptr(s2), stream_ptr())), "bn_bwd_reduce")
    return s1, s2


def _bn_bwd_apply(dz, z, x, act, mean, rstd, gamma, s1, s2, inv_n, batch_stats, dx, dres, acc_dres):
    vdz = view4(dz)
    vz = view4(z) if z is not None else _lib.null_view()
    vx = view4(x) if x is not None else _lib.null_view()
    vdx = view4(dx) if dx is not None else _lib.null_view()
    vdr = view4(dres) if dres is not None else _lib.null_view()
    check(profiler.launch("bn_bwd_apply", lambda: lib().pmoe_bn_bwd_apply(
        C.byref(vdz), C.byref(vz), C.byref(vx), dtype_code(dz), ACT[act], _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(gamma),
        _lib.ptr(s1), _lib.ptr(s2), float(inv_n), int(batch_stats), C.byref(vdx), C.byref(vdr), int(acc_dres),
        stream_ptr())), "bn_bwd_apply")


def _axpy(src, dst, alpha=1.0, bcast=None, accumulate=False):
    vs = view4(src) if src is not None else _lib.null_view()
    vd = view4(dst)
    check(profiler.launch("axpy", lambda: lib().pmoe_axpy(
        C.byref(vs), C.byref(vd), dtype_code(dst), float(alpha), _lib.ptr(bcast), 0 if bcast is None else bcast.stride(0),
        int(accumulate), stream_ptr())), "axpy")


def _accumulate(tape, act, g):
    """grads[act] += g (g already shaped like act.t)."""
    k = id(act)
    if k in tape.grads:
        _axpy(g, tape.grads[k], 1.0, None, True)
    else:
        tape.grads[k] = g


def _grad_buffer(tape, act):
    """Existing gradient buffer of `act` (to be accumulated into) or a fresh one; returns (tensor, existed)."""
    k = id(act)
    if k in tape.grads:
        return tape.grads[k], True
    g = torch.empty(act.t.shape, dtype=act.t.dtype, device=act.t.device)
    tape.grads[k] = g
    return g, False


# ------------------------------------------------------------------------------------------------ conv (+BN +act +residual)
def conv_op(tape, srcs, weight, bias=None, bn=None, act=None, residual=None, want_pool=False, ksize=3, layouts=None,
            taps=None, out=None, pool_out=None, pool_stride=0, tag=""):
    """conv / linear layer over the virtual concat of `srcs`, followed by BatchNorm (batch statistics when
    bn.training, folded running statistics otherwise), optional residual add and activation.
    Returns (Act z, pool_sum or None)."""
    dt = tape.dtype
    dev = srcs[0].t.device
    cout = weight.shape[0]
    cop = ops.cout_padded(cout)
    cstore = pad_ch(cout)
    lays = _layout_of(srcs, layouts)
    phys = [a.cpad for a in srcs]
    ck = ops.choose_ck(phys)
    if taps is None:
        taps = TAPS3 if ksize == 3 else [(0, 0)]
    pad = ksize // 2
    tap_off = [(r - pad, s - pad) for (r, s) in taps]
    segs = ops.conv_segments(tap_off, phys, ck)
    wp = _pack_fwd(weight, lays, taps, cop, dt)
    n, h, w, _ = srcs[0].t.shape
    cin = weight.shape[1]
    flops = 2.0 * n * h * w * cout * cin * len(taps)
    bn_train = bn is not None and bn.training
    rg_in = any(_rg(a) for a in srcs) or _any_rg([weight, bias]) or (bn is not None and _any_rg([bn.weight, bn.bias])) \
        or (residual is not None and _rg(residual))
    pool = None
    if want_pool:
        pool = pool_out if pool_out is not None else torch.zeros(n, cop, dtype=torch.float32, device=dev)
    raw = mean = rstd = gamma_p = scale = None
    if bn_train:
        raw = torch.empty(n, h, w, cstore, dtype=dt, device=dev)
        ssum = torch.zeros(cop, dtype=torch.float32, device=dev)
        ssq = torch.zeros(cop, dtype=torch.float32, device=dev)
        ops.conv([a.t for a in srcs], wp, segs, ck, raw, stat_sum=ssum, stat_sqsum=ssq, flops=flops, tag=tag)
        track = bn.track_running_stats and bn.running_mean is not None
        mom = 0.1 if bn.momentum is None else bn.momentum
        mean, rstd, scale, shift = nhwc.bn_finalize(ssum, ssq, n * h * w, cout, bn.weight.detach(), bn.bias.detach(), bn.eps, mom,
                                                    bn.running_mean if track else None, bn.running_var if track else None)
        if track and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        z_t = out if out is not None else torch.empty(n, h, w, cstore, dtype=dt, device=dev)
        nhwc.affine_act(raw, scale[:cstore], shift[:cstore], act, None if residual is None else residual.t, out=z_t)
        if want_pool:
            nhwc.channel_sums(z_t, out=pool)
        gamma_p = ops.pad_vec(bn.weight.detach(), cstore, 0.0)
    else:
        if bn is not None:
            scale = ops.pad_vec(bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps), cop, 0.0)
            shift = ops.pad_vec(bn.bias.detach().float() - bn.running_mean.float() * scale[:cout], cop, 0.0)
        else:
            shift = ops.pad_vec(bias.detach(), cop, 0.0) if bias is not None else None
        z_t = out if out is not None else torch.empty(n, h, w, cstore, dtype=dt, device=dev)
        ops.conv([a.t for a in srcs], wp, segs, ck, z_t, scale=scale, shift=shift, act=act,
                 residual=None if residual is None else residual.t, pool_sum=pool, pool_stride=pool_stride, flops=flops, tag=tag)
    z = _new_act(tape, z_t, cout, rg_in)
    if not (tape.save and rg_in):
        return z, pool

    def backward():
        dz = tape.grad_of(z)
        if dz is None:
            return
        z_saved = z_t if act not in (None, "none") else None
        dres = None
        acc_dres = False
        if residual is not None and _rg(residual):
            dres, acc_dres = _grad_buffer(tape, residual)
        dy = torch.empty(n, h, w, cstore, dtype=dt, device=dev)  # gradient w.r.t. the raw conv output
        if bn_train:
            s1, s2 = _bn_bwd_reduce(dz, z_saved, raw, act, mean, rstd, cstore)
            tape.add_pgrad(bn.weight, s2[:cout])
            tape.add_pgrad(bn.bias, s1[:cout])
            _bn_bwd_apply(dz, z_saved, raw, act, mean, rstd, gamma_p, s1, s2, 1.0 / (n * h * w), 1, dy, dres, acc_dres)
        else:
            if bn is not None and _any_rg([bn.weight, bn.bias]):
                raise NotImplementedError("pmoe_b200: gradients of BatchNorm affine parameters in eval mode are not supported")
            if bias is not None and bias.requires_grad:
                s1, _ = _bn_bwd_reduce(dz, z_saved, None, act, None, None, cstore)
                tape.add_pgrad(bias, s1[:cout])
            _bn_bwd_apply(dz, z_saved, None, act, None, None, None if scale is None else scale[:cstore].contiguous(), None, None,
                          0.0, 0, dy, dres, acc_dres)
        if weight.requires_grad:
            dwp = torch.zeros(cop, wp.shape[1], dtype=torch.float32, device=dev)
            ops.conv_wgrad([a.t for a in srcs], segs, ck, dy, dwp, flops=flops, tag="wgrad " + tag)
            tape.add_pgrad(weight, _unpack_wgrad(dwp, weight.shape, lays, len(taps)))
        # data gradients, one launch per physical source that needs them
        ck_d = ops.choose_ck([cstore])
        dtaps = [(-dh, -dw) for (dh, dw) in tap_off]
        cin_begin = 0
        for a, lay in zip(srcs, lays):
            nlog = sum(g[0] for g in lay)
            if _rg(a):
                wd = _pack_dgrad(weight, lay, cin_begin, taps, cstore, dt)
                dsegs = ops.conv_segments(dtaps, [cstore], ck_d)
                g, existed = _grad_buffer(tape, a)
                ops.conv([dy], wd, dsegs, ck_d, g, residual=g if existed else None,
                         flops=2.0 * n * h * w * cout * nlog * len(taps), tag="dgrad " + tag)
            cin_begin += nlog

    tape.record(backward)
    return z, pool


def maxpool_op(tape, x, k, stride, pad):
    y = nhwc.maxpool(x, k, stride, pad)
    ya = _new_act(tape, y.t, x.c, _rg(x))
    if tape.save and _rg(x):
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            g, existed = _grad_buffer(tape, x)
            vx, vdy, vdx = view4(x.t), view4(dy), view4(g)
            check(profiler.launch("maxpool_bwd", lambda: lib().pmoe_maxpool_bwd(
                C.byref(vx), C.byref(vdy), C.byref(vdx), dtype_code(g), k, stride, pad, int(existed), stream_ptr())), "maxpool_bwd")
        tape.record(backward)
    return ya


def conv_transpose_op(tape, up, x, tag=""):
    """nn.ConvTranspose2d(k=2, s=2): four 1x1 GEMMs into the pixel-shuffle views of the output."""
    dt, dev = tape.dtype, x.t.device
    weight, bias = up.weight, up.bias
    cin, cout = weight.shape[0], weight.shape[1]
    cop, cstore = ops.cout_padded(cout), pad_ch(cout)
    n, h, w, cp = x.t.shape
    ck = ops.choose_ck([cp])
    segs = ops.conv_segments([(0, 0)], [cp], ck)
    shift = ops.pad_vec(bias.detach(), cop, 0.0)
    out = torch.empty(n, 2 * h, 2 * w, cstore, dtype=dt, device=dev)
    wf = weight.detach().float()
    flops = 2.0 * n * h * w * cin * cout
    for a in range(2):
        for b in range(2):
            wp = ops.pack_conv_weight(wf[:, :, a, b].t().reshape(cout, cin, 1, 1), [x.c], [cp], [(0, 0)], cop, dt)
            ops.conv([x.t], wp, segs, ck, out[:, a::2, b::2, :], shift=shift, flops=flops, tag="convT " + tag)
    rg = _rg(x) or _any_rg([weight, bias])
    ya = _new_act(tape, out, cout, rg)
    if tape.save and rg:
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            if bias.requires_grad:
                s1, _ = _bn_bwd_reduce(dy, None, None, None, None, None, cstore)
                tape.add_pgrad(bias, s1[:cout])
            if weight.requires_grad:
                gw = torch.empty(cin, cout, 2, 2, dtype=torch.float32, device=dev)
                for a in range(2):
                    for b in range(2):
                        dwp = torch.zeros(cop, cp, dtype=torch.float32, device=dev)
                        ops.conv_wgrad([x.t], segs, ck, dy[:, a::2, b::2, :], dwp, flops=flops, tag="wgrad convT " + tag)
                        gw[:, :, a, b] = dwp[:cout, :cin].t()
                tape.add_pgrad(weight, gw)
            if _rg(x):
                # dx[n,h,w,ci] = sum_{a,b,co} dy[n,2h+a,2w+b,co] * W[ci,co,a,b]: 4 parity views of dy as 4 K segments
                views = [dy[:, a::2, b::2, :] for a in range(2) for b in range(2)]
                ck_d = ops.choose_ck([cstore])
                dsegs = [(q, 0, 0, 0, cstore // ck_d) for q in range(4)]
                blocks = []
                for a in range(2):
                    for b in range(2):
                        blk = wf[:, :, a, b]  # (cin, cout)
                        blocks.append(torch.nn.functional.pad(blk, (0, cstore - cout)))
                wd = torch.cat(blocks, 1)  # (cin, 4*cstore)
                rows = ops.cout_padded(cp)
                wd = torch.nn.functional.pad(wd, (0, 0, 0, rows - cin)).to(dt).contiguous()
                g, existed = _grad_buffer(tape, x)
                ops.conv(views, wd, dsegs, ck_d, g, residual=g if existed else None, flops=4 * flops, tag="dgrad convT " + tag)
        tape.record(backward)
    return ya


def eca_op(tape, eca_mod, x, layout=None, pool_in=None):
    """EfficientBlock: out = x * sigmoid(conv1d(mean_hw(x))). layout = (groups, logical, slot)."""
    w = eca_mod.conv.weight
    n, h, wd, cp = x.t.shape
    groups, gl, gs = layout if layout is not None else (1, x.c, cp)
    sums = pool_in if pool_in is not None else nhwc.channel_sums(x.t)
    gate = nhwc.eca_gate(sums, h * wd, w.detach(), groups, gl, gs)
    y = nhwc.scale_channels(x.t, gate)
    rg = _rg(x) or w.requires_grad
    ya = _new_act(tape, y, x.c, rg)
    if tape.save and rg:
        def backward():
            dy = tape.grad_of(ya)
            if dy is None:
                return
            dgate = torch.zeros(n, cp, dtype=torch.float32, device=dy.device)
            va, vb = view4(dy), view4(x.t)
            check(profiler.launch("prod_channel_sums", lambda: lib().pmoe_prod_channel_sums(
                C.byref(va), C.byref(vb), dtype_code(dy), dgate.data_ptr(), dgate.stride(0), stream_ptr())), "prod_channel_sums")
            dmean = torch.empty(n, cp, dtype=torch.float32, device=dy.device)
            dw = torch.zeros(w.numel(), dtype=torch.float32, device=dy.device)
            wf = w.detach().reshape(-1)
            check(profiler.launch("eca_gate_bwd", lambda: lib().pmoe_eca_gate_bwd(
                dgate.data_ptr(), dgate.stride(0), gate.data_ptr(), gate.stride(0), sums.data_ptr(), sums.stride(0), n,
                1.0 / float(h * wd), wf.data_ptr(), wf.numel(), groups, gl, gs, dmean.data_ptr(), dmean.stride(0),
                dw.data_ptr(), stream_ptr())), "eca_gate_bwd")
            tape.add_pgrad(w, dw)
            if _rg(x):
                g, existed = _grad_buffer(tape, x)
                vd, vg = view4(dy), view4(g)
                check(profiler.launch("eca_bwd_apply", lambda: lib().pmoe_eca_bwd_apply(
                    C.byref(vd), dtype_code(dy), gate.data_ptr(), gate.stride(0), dmean.data_ptr(), dmean.stride(0), C.byref(vg),
                    int(existed), stream_ptr())), "eca_bwd_apply")
        tape.record(backward)
    return ya


# ------------------------------------------------------------------------------------------------ module runners
def conv3_block(tape, seq, srcs, want_pool=False, tag=""):
    y, _ = conv_op(tape, srcs, seq[0].weight, None, seq[1], "relu", tag=tag + ".0")
    return conv_op(tape, [y], seq[3].weight, None, seq[4], "relu", want_pool=want_pool, tag=tag + ".3")


def unet(tape, net, x, out=None, out_pool=None, pool_stride=0, want_inter=False, tag="unet"):
    """UNet.forward (unet.py:50-95). Returns (Act logits, pooled bottleneck (N,512) fp32 or None)."""
    n, h, w, _ = x.t.shape
    if h % 16 or w % 16:
        raise RuntimeError("pmoe_b200 UNet needs H and W divisible by 16 (got %dx%d)" % (h, w))
    x1, _ = conv3_block(tape, net.dwn_1, [x], tag=tag + ".dwn_1")
    x2, _ = conv3_block(tape, net.dwn_2, [maxpool_op(tape, x1, 2, 2, 0)], tag=tag + ".dwn_2")
    x3, _ = conv3_block(tape, net.dwn_3, [maxpool_op(tape, x2, 2, 2, 0)], tag=tag + ".dwn_3")
    x4, _ = conv3_block(tape, net.dwn_4, [maxpool_op(tape, x3, 2, 2, 0)], tag=tag + ".dwn_4")
    x5, pool5 = conv3_block(tape, net.dwn_5, [maxpool_op(tape, x4, 2, 2, 0)], want_pool=want_inter, tag=tag + ".dwn_5")
    y = x5
    for i, (up, fwd, skip) in enumerate(((net.up_1, net.up_forw_1, x4), (net.up_2, net.up_forw_2, x3),
                                         (net.up_3, net.up_forw_3, x2), (net.up_4, net.up_forw_4, x1)), start=1):
        u = conv_transpose_op(tape, up, y, tag=tag + ".up_%d" % i)
        y, _ = conv3_block(tape, fwd, [skip, u], tag=tag + ".up_forw_%d" % i)
    logits, _ = conv_op(tape, [y], net.out.weight, net.out.bias, None, None, ksize=1, out=out, want_pool=out_pool is not None,
                        pool_out=out_pool, pool_stride=pool_stride, tag=tag + ".out")
    inter = None
    if want_inter:
        inter = InterRepr(tape, x5, pool5)
    return logits, inter


class InterRepr:
    """adaptive_avg_pool2d(x_5, 1).flatten(1) (unet.py:89-92) with its backward into x_5."""

    def __init__(self, tape, x5, pool):
        self.tape, self.x5 = tape, x5
        self.hw = x5.t.shape[1] * x5.t.shape[2]
        self.value = pool[:, :x5.c] / float(self.hw)

    def backward(self, g):
        """g: (N, C) fp32 gradient of the pooled vector."""
        if not _rg(self.x5):
            return
        gb = torch.zeros(self.x5.t.shape[0], self.x5.cpad, dtype=torch.float32, device=g.device)
        gb[:, :self.x5.c] = g.float() / float(self.hw)
        buf, existed = _grad_buffer(self.tape, self.x5)
        _axpy(None, buf, 1.0, gb, existed)


def eca_conv_block(tape, blk, x, layout=None, pool_in=None, tag="eca_block"):
    """EfficientConvBlock (basics.py:80-135). layout = (groups, logical, slot) of x's channel axis."""
    xs = eca_op(tape, blk.layer1.eca1, x, layout, pool_in)
    lay = None if layout is None else [[(layout[1], layout[2])] * layout[0]]
    c1, pool64 = conv_op(tape, [xs], blk.layer1.conv1[0].weight, None, blk.layer1.conv1[1], "relu", want_pool=True,
                         layouts=lay, tag=tag + ".conv1")
    c1s = eca_op(tape, blk.layer2.eca2, c1, None, pool64)
    y, _ = conv_op(tape, [c1s], blk.layer2.conv2[0].weight, None, blk.layer2.conv2[1], "relu", tag=tag + ".conv2")
    return y


# ------------------------------------------------------------------------------------------------ autograd bridge
class TapeFunction(torch.autograd.Function):
    """forward(runner, *params): runner(tape) -> (list of output tensors, seed_fn). seed_fn(grad_outputs)
    installs the output gradients into the tape; backward then replays the tape and returns the parameter
    gradients accumulated in it."""

    @staticmethod
    def forward(ctx, runner, *params):
        tape = Tape(config.act_dtype(), save=any(p.requires_grad for p in params))
        outs, seed = runner(tape)
        ctx.tape, ctx.seed, ctx.plist = tape, seed, params
        ctx.mark_non_differentiable(*[o for o in outs if not o.is_floating_point()])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        tape = ctx.tape
        ctx.seed(tape, gouts)
        tape.backward()
        grads = tuple(tape.pgrads.get(id(p)) if p.requires_grad else None for p in ctx.plist)
        ctx.tape = None
        return (None,) + grads


def run(module, runner):
    """Execute `runner(tape)` for `module`: through autograd when gradients are wanted, directly otherwise."""
    params = [p for p in module.parameters()]
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        outs = TapeFunction.apply(runner, *params)
    else:
        tape = Tape(config.act_dtype(), save=False)
        outs, _ = runner(tape)
    return outs


def _seed_nchw(act):
    """seed function for an NHWC output Act that was exported as an fp32 NCHW tensor."""
    def seed(tape, g):
        if g is None or not _rg(act):
            return
        buf = torch.empty(act.t.shape, dtype=act.t.dtype, device=act.t.device)
        nhwc.from_nchw(g.contiguous(), out=buf)
        _accumulate(tape, act, buf)
    return seed


def unet_module_forward(net, image):
    """UNet.forward for the module API: image fp32 NCHW -> logits fp32 NCHW (and pooled bottleneck)."""
    def runner(tape):
        x = nhwc.from_nchw(image, dtype=tape.dtype)
        x.rg = False
        logits, inter = unet(tape, net, x, want_inter=net.inter_repr)
        out = nhwc.to_nchw(logits.t, logits.c)
        seed_logits = _seed_nchw(logits)
        if net.inter_repr:
            def seed(tp, gouts):
                if gouts[0] is not None:
                    inter.backward(gouts[0])
                seed_logits(tp, gouts[1])
            return [inter.value, out], seed
        return [out], (lambda tp, gouts: seed_logits(tp, gouts[0]))
    outs = run(net, runner)
    return (outs[0], outs[1]) if net.inter_repr else outs[0]


def punet_module_forward(net, images):
    """PredictiveUnet.forward (punet.py:75-120) in the general path: every U-Net call is separate (train-mode
    BatchNorm statistics are per call, as in the reference's Python loop)."""
    B, T, Cin, H, W = images.shape
    P, Fu = net.n_past_frames, net.n_future_frames
    ncls = net.unet.out.weight.shape[0]
    slot = pad_ch(ncls)

    def runner(tape):
        dev = images.device
        nslots = P + max(Fu, 0)
        ring = torch.zeros(B, H, W, nslots * slot, dtype=tape.dtype, device=dev)
        pools = torch.zeros(B, nslots * slot, dtype=torch.float32, device=dev)
        masks = []
        inter = None
        for t in range(P):
            x = nhwc.from_nchw(images[:, t], dtype=tape.dtype)
            x.rg = False
            m, it = unet(tape, net.unet, x, out=ring[..., t * slot:(t + 1) * slot], out_pool=pools[:, t * slot:],
                         pool_stride=nslots * slot, want_inter=(net.unet_inter_repr and Fu == 0 and t == P - 1), tag="unet")
            masks.append(m)
            inter = it
        if Fu == 0:
            if net.unet_inter_repr:
                return [inter.value], (lambda tp, g: inter.backward(g[0]) if g[0] is not None else None)
            return [nhwc.to_nchw(masks[-1].t, ncls)], (lambda tp, g: _seed_nchw(masks[-1])(tp, g[0]))
        futures = []
        for f in range(Fu):
            window = _new_act(tape, ring[..., f * slot:(f + P) * slot], P * ncls, any(_rg(m) for m in masks[f:f + P]))
            window_parts = masks[f:f + P]
            if tape.save and _rg(window):
                def split_grad(window=window, parts=window_parts):
                    g = tape.grad_of(window)
                    if g is None:
                        return
                    for i, part in enumerate(parts):
                        if _rg(part):
                            _accumulate_copy(tape, part, g[..., i * slot:(i + 1) * slot])
                tape.record(split_grad)
            e = eca_conv_block(tape, net.entry_block, window, (P, ncls, slot), pools[:, f * slot:(f + P) * slot], tag="entry")
            m, inter = unet(tape, net.pred_unet, e, out=ring[..., (P + f) * slot:(P + f + 1) * slot],
                            out_pool=pools[:, (P + f) * slot:], pool_stride=nslots * slot, want_inter=net.inter_repr,
                            tag="pred_unet")
            masks.append(m)
            futures.append(m)
        if net.inter_repr:
            return [inter.value], (lambda tp, g: inter.backward(g[0]) if g[0] is not None else None)
        out = torch.empty(B, Fu, ncls, H, W, dtype=torch.float32, device=dev)
        for f, m in enumerate(futures):
            nhwc.to_nchw(m.t, ncls, out=out[:, f])

        def seed(tp, gouts):
            g = gouts[0]
            if g is None:
                return
            for f, m in enumerate(futures):
                _seed_nchw(m)(tp, g[:, f])
        return [out], seed

    return run(net, runner)[0]


def _accumulate_copy(tape, act, gview):
    """grads[act] += gview where gview is a strided view with act's geometry."""
    k = id(act)
    if k in tape.grads:
        _axpy(gview, tape.grads[k], 1.0, None, True)
    else:
        buf = torch.empty(act.t.shape[0], act.t.shape[1], act.t.shape[2], act.t.shape[3], dtype=gview.dtype, device=gview.device)
        _axpy(gview, buf, 1.0, None, False)
        tape.grads[k] = buf


def nhwc_module_forward(module, x_nchw, body):
    """Run `body(tape, Act) -> Act` for a stand-alone block called through the NCHW module API."""
    def runner(tape):
        x = nhwc.from_nchw(x_nchw, dtype=tape.dtype)
        y = body(tape, x)
        return [nhwc.to_nchw(y.t, y.c)], (lambda tp, g: _seed_nchw(y)(tp, g[0]))
    return run(module, runner)[0]
