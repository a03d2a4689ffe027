"""Host-side wrappers over the C-ABI: weight packing, segment lists and launch helpers.

Everything here prepares descriptors; all arithmetic happens in libpmoe_b200.so on the GPU.
Activation tensors are NHWC torch tensors (N,H,W,Cpad) whose channel count is padded to a
multiple of 16 with zeros (`pad_ch`).
"""
import ctypes as C

import torch

from . import _lib, profiler
from ._lib import ACT


def pad_ch(c):
    return (int(c) + 15) // 16 * 16


def choose_ck(cpads):
    for ck in (64, 32, 16):
        if all(c % ck == 0 for c in cpads):
            return ck
    raise ValueError("channel counts must be multiples of 16: %r" % (cpads,))


def cout_padded(cout):
    """Rows of the packed weight: a multiple of the N tile the kernel will pick."""
    c = pad_ch(cout)
    if c > 256:
        c = (c + 255) // 256 * 256
    elif c > 128:
        c = 256
    elif c > 64:
        c = 128
    elif c > 32:
        c = 64
    return c


def pack_conv_weight(w, src_channels, src_cpads, taps, cout_pad, dtype=torch.bfloat16):
    """w: (Cout, sum(src_channels), R, S). Returns [cout_pad, ktot] with K enumerating
    (tap, source, padded channel) — the order `conv_segments` emits."""
    cout = w.shape[0]
    cols = []
    for (r, s) in taps:
        c_begin = 0
        for c, cp in zip(src_channels, src_cpads):
            blk = w[:, c_begin:c_begin + c, r, s]
            if cp > c:
                blk = torch.nn.functional.pad(blk, (0, cp - c))
            cols.append(blk)
            c_begin += c
    wp = torch.cat(cols, dim=1)
    if cout_pad > cout:
        wp = torch.nn.functional.pad(wp, (0, 0, 0, cout_pad - cout))
    return wp.to(dtype).contiguous()


def conv_segments(taps_dhdw, src_cpads, ck):
    """[(src, dh, dw, c0, nchunks)] for each tap x source."""
    segs = []
    for (dh, dw) in taps_dhdw:
        for i, cp in enumerate(src_cpads):
            segs.append((i, dh, dw, 0, cp // ck))
    return segs


def pack_convT_weight(up, cp, cs, dtype=None):
    """ConvTranspose2d(k=2,s=2) weight (Cin, Cout, 2, 2) + bias -> ([4*cs, cp] packed weight, [4*cs] shift): row block
    q = 2a+b carries W[:, :, a, b]^T (the GEMM of output parity (a, b)); the bias repeats per block."""
    from . import config
    dtype = config.act_dtype() if dtype is None else dtype
    wf = up.weight.detach().float()
    cin, cout = wf.shape[0], wf.shape[1]
    wp = torch.zeros(4 * cs, cp, dtype=torch.float32, device=wf.device)
    shift = torch.zeros(4 * cs, dtype=torch.float32, device=wf.device)
    for q, (a, b) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        wp[q * cs:q * cs + cout, :cin] = wf[:, :, a, b].t()
        if up.bias is not None:
            shift[q * cs:q * cs + cout] = up.bias.detach().float()
    return wp.to(dtype).contiguous(), shift


TAPS3 = [(r, s) for r in range(3) for s in range(3)]


def pad_vec(v, n, fill=0.0):
    out = torch.full((n,), fill, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v.float()
    return out


def _fill_desc(srcs, wpack, segs, ck, out, scale, shift, act, residual, stat_sum, stat_sqsum, pool_sum, src_channels,
               pool_stride, dtype, out_extra=None, out_cols=0, pool2_out=None, nchw_out=None):
    d = _lib.ConvTc()
    assert 1 <= len(srcs) <= _lib.MAX_SRC and 1 <= len(segs) <= _lib.MAX_SEG, (len(srcs), len(segs))
    d.n_src = len(srcs)
    for i, s in enumerate(srcs):
        _lib.require_cuda(s, "conv source")
        assert s.dtype == dtype, (s.dtype, dtype)
        d.src[i] = _lib.view4(s, None if src_channels is None else src_channels[i])
    d.n_seg = len(segs)
    for i, (src, dh, dw, c0, nch) in enumerate(segs):
        d.seg[i].src, d.seg[i].dh, d.seg[i].dw, d.seg[i].c0, d.seg[i].nchunks = src, dh, dw, c0, nch
    d.ck = ck
    assert wpack.is_contiguous()
    if wpack.dim() == 3:  # (n_img, cout_pad, ktot): one weight set per image (gate_weights)
        d.wpack_img_stride = wpack.stride(0)
        wpack = wpack[0]
    d.ktot = wpack.shape[1]
    d.cout_pad = wpack.shape[0]
    d.wpack = wpack.data_ptr()
    assert out.dtype == dtype, (out.dtype, dtype)
    d.out = _lib.view4(out)
    d.scale = None if scale is None else scale.data_ptr()
    d.shift = None if shift is None else shift.data_ptr()
    if shift is not None and shift.dim() == 2:  # (n_img, cout_pad): one bias row per image / expert
        d.shift_img_stride = shift.stride(0)
    d.act = ACT[act]
    d.residual = _lib.view4(residual) if residual is not None else _lib.null_view()
    d.stat_sum = None if stat_sum is None else stat_sum.data_ptr()
    d.stat_sqsum = None if stat_sqsum is None else stat_sqsum.data_ptr()
    d.pool_sum = None if pool_sum is None else pool_sum.data_ptr()
    d.pool_stride = int(pool_stride) if pool_stride else (pool_sum.stride(0) if pool_sum is not None and pool_sum.dim() == 2 else 0)
    if out_extra:
        assert len(out_extra) <= 3
        d.n_out_extra = len(out_extra)
        d.out_cols = int(out_cols)
        for i, t in enumerate(out_extra):
            assert t.dtype == dtype
            d.out_extra[i] = _lib.view4(t)
    if pool2_out is not None:
        assert pool2_out.dtype == dtype
        d.pool2_out = _lib.view4(pool2_out)
    if nchw_out is not None:  # fp32 (N, C, H, W) tensor/view
        assert nchw_out.dtype == torch.float32 and nchw_out.dim() == 4
        d.nchw_out = nchw_out.data_ptr()
        d.nchw_sn, d.nchw_sc, d.nchw_sh, d.nchw_sw = nchw_out.stride()
        d.nchw_c = nchw_out.shape[1]
    return d


def conv_tc(srcs, wpack, segs, ck, out, scale=None, shift=None, act=None, residual=None, stat_sum=None,
            stat_sqsum=None, pool_sum=None, src_channels=None, pool_stride=0, flops=0.0, tag="", out_extra=None, out_cols=0,
            pool2_out=None, nchw_out=None):
    """Launch the tcgen05 implicit-GEMM kernel. srcs: list of NHWC bf16 tensors (views allowed);
    out: NHWC bf16 tensor/view; wpack: [cout_pad, ktot] bf16. pool2_out: optional fused MaxPool2d(2,2) of the output;
    nchw_out: optional fp32 NCHW copy of the first channels."""
    assert wpack.dtype == torch.bfloat16
    d = _fill_desc(srcs, wpack, segs, ck, out, scale, shift, act, residual, stat_sum, stat_sqsum, pool_sum, src_channels,
                   pool_stride, torch.bfloat16, out_extra, out_cols, pool2_out, nchw_out)
    fn = _lib.lib().pmoe_conv_tc
    sp = _lib.stream_ptr()
    _lib.check(profiler.launch("conv_tc", lambda: fn(C.byref(d), sp), flops, 0.0, tag), "conv_tc")
    return out


def conv_simt(srcs, wpack, segs, ck, out, scale=None, shift=None, act=None, residual=None, stat_sum=None,
              stat_sqsum=None, pool_sum=None, src_channels=None, pool_stride=0, flops=0.0, tag="", out_extra=None, out_cols=0,
              pool2_out=None, nchw_out=None):
    """CUDA-core twin of conv_tc (fp32 or bf16 storage, fp32 FMA accumulation)."""
    dt = out.dtype
    assert wpack.dtype == dt
    assert pool2_out is None and nchw_out is None, "fused pool / NCHW copy are tensor-core-kernel features"
    if out_extra:
        # multi-view outputs are a tensor-core-kernel feature: one launch per view here
        cols = int(out_cols)
        for q, view in enumerate([out] + list(out_extra)):
            conv_simt(srcs, wpack[q * cols:(q + 1) * cols], segs, ck, view, None if scale is None else scale[q * cols:(q + 1) * cols],
                      None if shift is None else shift[q * cols:(q + 1) * cols], act, flops=flops / (len(out_extra) + 1), tag=tag)
        return out
    det_pool = pool_sum is not None and dt == torch.float32   # parity mode: per-image sums by the order-independent reduction kernel
    d = _fill_desc(srcs, wpack, segs, ck, out, scale, shift, act, residual, stat_sum, stat_sqsum, None if det_pool else pool_sum,
                   src_channels, pool_stride, dt)
    fn = _lib.lib().pmoe_conv_simt
    sp = _lib.stream_ptr()
    code = _lib.BF16 if dt == torch.bfloat16 else _lib.F32
    _lib.check(profiler.launch("conv_simt", lambda: fn(C.byref(d), code, sp), flops, 0.0, tag), "conv_simt")
    if det_pool:
        v = _lib.view4(out)
        stride = int(pool_stride) if pool_stride else (pool_sum.stride(0) if pool_sum.dim() == 2 else out.shape[3])
        _lib.check(profiler.launch("channel_sums", lambda: _lib.lib().pmoe_channel_sums(C.byref(v), code, pool_sum.data_ptr(), stride, sp)),
                   "channel_sums")
    return out


def gate_weights(wpack, gate, cphys):
    """(cout_pad, ktot) bf16 packed weights x (N, >= cphys) fp32 gate -> (N, cout_pad, ktot) per-image weights."""
    n = gate.shape[0]
    out = torch.empty(n, wpack.shape[0], wpack.shape[1], dtype=wpack.dtype, device=wpack.device)
    _lib.check(profiler.launch("gate_weights", lambda: _lib.lib().pmoe_gate_weights(
        wpack.data_ptr(), gate.data_ptr(), gate.stride(0), n, wpack.shape[0], wpack.shape[1], cphys, out.data_ptr(), _lib.stream_ptr())),
        "gate_weights")
    return out


def conv(srcs, wpack, segs, ck, out, **kw):
    """bf16 -> tensor cores, fp32 -> CUDA cores. Same arguments as conv_tc."""
    from . import config
    if out.dtype == torch.bfloat16 and not config.FORCE_SIMT:
        return conv_tc(srcs, wpack, segs, ck, out, **kw)
    return conv_simt(srcs, wpack, segs, ck, out, **kw)


def conv_wgrad(srcs, segs, ck, dy, dwpack, flops=0.0, tag=""):
    """dwpack[cout_pad, ktot] (fp32) += sum_pixels dy[p, co] * x_k[p] in the packed K order of the forward."""
    dt = dy.dtype
    from . import config
    if (dt == torch.bfloat16 and not config.FORCE_SIMT and dy.shape[3] < 64 and (dy.shape[0] * dy.shape[1] * dy.shape[2] >= 8192 or dwpack.dim() == 3)
            and dwpack.shape[-2] <= 64):
        # Skinny outputs (the 512->1 / 512->4 expert-head Linears over 64k rows): the tensor-core kernel wants >= 64 gradient
        # channels, and the CUDA-core fallback took 183 us per expert and layer for a 67 MB read. Zero-pad dy to 64 channels
        # (8 MB at 64k rows) and run the tensor-core kernel; only the first rows of its result are real.
        n, h, w, c = dy.shape
        dyp = torch.zeros(n, h, w, 64, dtype=dt, device=dy.device)
        vs, vd = _lib.view4(dy), _lib.view4(dyp[..., :c])
        _lib.check(profiler.launch("axpy", lambda: _lib.lib().pmoe_axpy(C.byref(vs), C.byref(vd), _lib.BF16, 1.0, None, 0, 0,
                                                                        _lib.stream_ptr()), io=(dy, dy)), "axpy")
        grouped = dwpack.dim() == 3   # (K, cout_pad, ktot): one gradient per image / expert
        wide = torch.zeros(((dwpack.shape[0],) if grouped else ()) + (64, dwpack.shape[-1]), dtype=torch.float32, device=dy.device)
        dw = _fill_desc(srcs, wide, segs, ck, dyp, None, None, None, None, None, None, None, None, 0, dt)
        tc = _lib.lib().pmoe_conv_wgrad_tc
        sp = _lib.stream_ptr()
        rc = profiler.launch("conv_wgrad_tc", lambda: tc(C.byref(dw), wide.data_ptr(), sp), flops, 0.0, tag)
        if rc == 0:
            dwpack += wide[..., :dwpack.shape[-2], :]
            return dwpack
        if rc != -2:
            _lib.check(rc, "conv_wgrad_tc")
        profiler.uncount()
    if dwpack.dim() == 3 and (dt != torch.bfloat16 or config.FORCE_SIMT):
        raise RuntimeError("pmoe_b200 conv_wgrad: the grouped (per-image) weight gradient is a tensor-core kernel feature")
    d = _fill_desc(srcs, dwpack, segs, ck, dy, None, None, None, None, None, None, None, None, 0, dt)
    assert dwpack.dtype == torch.float32
    sp = _lib.stream_ptr()
    if dt == torch.bfloat16 and not config.FORCE_SIMT:
        # tensor-core kernel first; PMOE_ERR_UNSUPPORTED (-2) means "shape not covered", nothing was launched
        tc = _lib.lib().pmoe_conv_wgrad_tc
        rc = profiler.launch("conv_wgrad_tc", lambda: tc(C.byref(d), dwpack.data_ptr(), sp), flops, 0.0, tag)
        if rc == 0:
            return dwpack
        if rc != -2 or dwpack.dim() == 3:
            _lib.check(rc, "conv_wgrad_tc")
        profiler.uncount()
    fn = _lib.lib().pmoe_conv_wgrad_simt
    code = _lib.BF16 if dt == torch.bfloat16 else _lib.F32
    _lib.check(profiler.launch("conv_wgrad", lambda: fn(C.byref(d), code, dwpack.data_ptr(), sp), flops, 0.0, tag), "conv_wgrad")
    return dwpack
