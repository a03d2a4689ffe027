"""NHWC activation handles and thin launch wrappers over the memory-bound C-ABI kernels."""
import ctypes as C

import torch

from . import _lib, profiler
from ._lib import ACT, BF16, F32, lib, check, stream_ptr, view4
from .ops import pad_ch


def dtype_code(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError("pmoe_b200 kernels take bf16 or fp32 activations, got %s" % t.dtype)


class Act:
    """An NHWC activation: tensor (N,H,W,Cpad) with `c` logical channels; channels [c, Cpad) are zero."""
    __slots__ = ("t", "c", "rg", "stats", "bn_relu", "bn_ctx", "gate_grad")

    def __init__(self, t, c, rg=False):
        self.t, self.c, self.rg = t, c, rg
        self.gate_grad = None  # ECA of a network input: callback(d gate (N, Cpad) fp32) fed by the consuming conv's per-image wgrad
        self.bn_relu = False  # t = relu(BatchNorm(raw conv output)) with batch statistics and nothing added (set by train.conv_op)
        self.bn_ctx = None  # with bn_relu: (bn, raw, mean, rstd, (scale, shift), gamma, cout, cstore) of that BatchNorm (train.conv_op)
        self.stats = None   # (sum, sum of squares[, count > 0]) per channel, fp64, when the producing kernel already reduced them

    @property
    def shape(self):
        return self.t.shape

    @property
    def cpad(self):
        return self.t.shape[3]

    def slice_batch(self, a, b):
        return Act(self.t[a:b], self.c)


def empty_act(n, h, w, c, dtype, device, cpad=None):
    return Act(torch.empty(n, h, w, pad_ch(c) if cpad is None else cpad, dtype=dtype, device=device), c)


def from_nchw(x, dtype=torch.bfloat16, out=None):
    """fp32 (N,C,H,W) strided tensor -> NHWC Act (channels zero-padded to 16)."""
    _lib.require_cuda(x, "input")
    if x.dtype != torch.float32:
        x = x.float()
    n, c, h, w = x.shape
    if out is None:
        out = torch.empty(n, h, w, pad_ch(c), dtype=dtype, device=x.device)
    v = view4(out)
    check(profiler.launch("nchw_to_nhwc", lambda: lib().pmoe_nchw_to_nhwc(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), x.stride(3), c, C.byref(v),
                                  dtype_code(out), stream_ptr()), io=(x, out)), "nchw_to_nhwc")
    return Act(out, c)


def to_nchw(t, c, out=None):
    """NHWC tensor/view -> fp32 (N,c,H,W) (written into `out` if given, any strides)."""
    n, h, w, _ = t.shape
    if out is None:
        out = torch.empty(n, c, h, w, dtype=torch.float32, device=t.device)
    v = view4(t)
    check(profiler.launch("nhwc_to_nchw", lambda: lib().pmoe_nhwc_to_nchw(C.byref(v), dtype_code(t), c, out.data_ptr(), out.stride(0), out.stride(1), out.stride(2),
                                  out.stride(3), stream_ptr()), io=(t[..., :c], out)), "nhwc_to_nchw")
    return out


def maxpool(x, k, stride, pad, scale=None, shift=None, relu=False, want_idx=False):
    """MaxPool2d on an NHWC Act. want_idx: also return the uint8 argmax codes (N,OH,OW,Cpad) for the backward."""
    n, h, w, cp = x.t.shape
    oh = (h + 2 * pad - k) // stride + 1
    ow = (w + 2 * pad - k) // stride + 1
    out = torch.empty(n, oh, ow, cp, dtype=x.t.dtype, device=x.t.device)
    vs, vd = view4(x.t), view4(out)
    if want_idx:
        idx = torch.empty(n, oh, ow, cp, dtype=torch.uint8, device=x.t.device)
        check(profiler.launch("maxpool", lambda: lib().pmoe_maxpool_idx(C.byref(vs), C.byref(vd), dtype_code(out), k, stride, pad,
                                                                        idx.data_ptr(), stream_ptr()), io=(x.t, out, idx)), "maxpool_idx")
        return Act(out, x.c), idx
    check(profiler.launch("maxpool", lambda: lib().pmoe_maxpool(C.byref(vs), C.byref(vd), dtype_code(out), k, stride, pad, _lib.ptr(scale), _lib.ptr(shift),
                             int(relu), stream_ptr()), io=(x.t, out)), "maxpool")
    return Act(out, x.c)


def channel_sums(t, out=None):
    n, _, _, cp = t.shape
    if out is None:
        out = torch.zeros(n, cp, dtype=torch.float32, device=t.device)
    v = view4(t)
    check(profiler.launch("channel_sums", lambda: lib().pmoe_channel_sums(C.byref(v), dtype_code(t), out.data_ptr(), out.stride(0), stream_ptr()), io=(t,)), "channel_sums")
    return out


def eca_gate(pool_sum, count, w, groups, group_c, group_stride):
    """pool_sum: fp32 (N, >=groups*group_stride) view with unit inner stride. Returns gate (N, groups*group_stride)."""
    n = pool_sum.shape[0]
    gate = torch.empty(n, groups * group_stride, dtype=torch.float32, device=pool_sum.device)
    wf = w.reshape(-1)
    check(profiler.launch("eca_gate", lambda: lib().pmoe_eca_gate(pool_sum.data_ptr(), pool_sum.stride(0), n, 1.0 / float(count), wf.data_ptr(), wf.numel(), groups,
                              group_c, group_stride, gate.data_ptr(), gate.stride(0), stream_ptr())), "eca_gate")
    return gate


def scale_channels(t, gate, out=None):
    if out is None:
        out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
    vs, vd = view4(t), view4(out)
    check(profiler.launch("scale_channels", lambda: lib().pmoe_scale_channels(C.byref(vs), C.byref(vd), dtype_code(t), gate.data_ptr(), gate.stride(0), stream_ptr()), io=(t, out)), "scale_channels")
    return out


def bn_finalize(stat_sum, stat_sq, count, c, gamma, beta, eps, momentum, running_mean, running_var):
    cp = stat_sum.numel()
    dev = stat_sum.device
    mean = torch.empty(cp, dtype=torch.float32, device=dev)
    rstd = torch.empty(cp, dtype=torch.float32, device=dev)
    scale = torch.empty(cp, dtype=torch.float32, device=dev)
    shift = torch.empty(cp, dtype=torch.float32, device=dev)
    check(profiler.launch("bn_finalize", lambda: lib().pmoe_bn_finalize(stat_sum.data_ptr(), stat_sq.data_ptr(), float(count), c, cp, _lib.ptr(gamma), _lib.ptr(beta),
                                 eps, momentum, _lib.ptr(running_mean), _lib.ptr(running_var), mean.data_ptr(), rstd.data_ptr(),
                                 scale.data_ptr(), shift.data_ptr(), stream_ptr())), "bn_finalize")
    return mean, rstd, scale, shift


def affine_act(t, scale, shift, act=None, residual=None, out=None):
    if out is None:
        out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
    vs, vd = view4(t), view4(out)
    vr = view4(residual) if residual is not None else _lib.null_view()
    check(profiler.launch("affine_act", lambda: lib().pmoe_affine_act(C.byref(vs), C.byref(vd), dtype_code(t), scale.data_ptr(), shift.data_ptr(), C.byref(vr),
                                ACT[act], stream_ptr()), io=(t, out, residual)), "affine_act")
    return out


def affine_act_stats(t, scale, shift, act, out, pool=None, pool_stride=0, out_stats=None):
    """affine_act fused with the statistics of the stored output a following layer needs: `pool` (N, >=C) fp32 per-image
    channel sums and/or `out_stats` = (sum, sqsum) fp64 per channel (both accumulate into zeroed buffers). Dense bf16 only:
    returns False (nothing launched) when the tensors do not qualify, and the caller runs the separate kernels."""
    if t.dtype != torch.bfloat16 or not t.is_contiguous() or not out.is_contiguous() or act not in (None, "none", "relu") \
            or t.shape[3] // 8 > 256:
        return False
    vs, vd = view4(t), view4(out)
    if out_stats is not None and len(out_stats) > 2:
        if pool is not None or act != "relu":
            return False
        check(profiler.launch("affine_act", lambda: lib().pmoe_affine_relu_stats_pos(
            C.byref(vs), C.byref(vd), scale.data_ptr(), shift.data_ptr(), out_stats[0].data_ptr(), out_stats[1].data_ptr(),
            out_stats[2].data_ptr(), stream_ptr()), io=(t, out)), "affine_relu_stats_pos")
        return True
    check(profiler.launch("affine_act", lambda: lib().pmoe_affine_act_stats(
        C.byref(vs), C.byref(vd), dtype_code(t), scale.data_ptr(), shift.data_ptr(), ACT[act], _lib.ptr(pool),
        (pool.stride(0) if pool_stride == 0 else pool_stride) if pool is not None else 0,
        _lib.ptr(out_stats[0] if out_stats else None), _lib.ptr(out_stats[1] if out_stats else None), stream_ptr()),
        io=(t, out)), "affine_act_stats")
    return True
