"""Loss functions with the reference's names and signatures (reference: PMoE/trainer/loss.py), computed by
fused CUDA kernels: one pass forward (+ the analytic gradient), one pass backward."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib, profiler
from ._lib import check, lib, stream_ptr


def _f32c(t):
    return t.detach().contiguous().float()


class _MoeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, mean, std, speed_pred, actions_gt, speed_gt, c0, c1):
        _lib.require_cuda(probs, "probs")
        B, K = probs.shape
        p, m, s = _f32c(probs), _f32c(mean), _f32c(std)
        sp = _f32c(speed_pred).reshape(B, -1)
        speed_k = sp.shape[1]
        ag, sg = _f32c(actions_gt).reshape(B, 2), _f32c(speed_gt).reshape(B)
        out = torch.zeros(3, dtype=torch.float32, device=probs.device)
        dp, dm, ds, dsp = torch.empty_like(p), torch.empty_like(m), torch.empty_like(s), torch.empty_like(sp)
        check(profiler.launch("moe_loss", lambda: lib().pmoe_moe_loss(
            p.data_ptr(), m.data_ptr(), s.data_ptr(), sp.data_ptr(), speed_k, ag.data_ptr(), sg.data_ptr(), B, K, float(c0),
            float(c1), out.data_ptr(), dp.data_ptr(), dm.data_ptr(), ds.data_ptr(), dsp.data_ptr(), None, stream_ptr())), "moe_loss")
        ctx.save_for_backward(dp, dm, ds, dsp.reshape(speed_pred.shape))
        return out[0]

    @staticmethod
    def backward(ctx, g):
        dp, dm, ds, dsp = ctx.saved_tensors
        return g * dp, g * dm, g * ds, g * dsp, None, None, None, None


def moe_loss(action_dists, speed_pred, actions_gt, speed_gt, loss_coefs):
    """NLL of the action mixture + speed MSE (loss.py:121-132), incl. the in-place unsqueeze_ of speed_gt."""
    probs = action_dists.mixture_distribution.probs
    base = action_dists.component_distribution.base_dist
    if len(speed_pred.shape) > 2:
        speed_gt.unsqueeze_(1)
    return _MoeLossFn.apply(probs, base.loc, base.scale, speed_pred, actions_gt, speed_gt, loss_coefs[0], loss_coefs[1])


class _L1MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, is_mse, coef):
        _lib.require_cuda(a, "prediction")
        af, bf = _f32c(a), _f32c(b).expand_as(a).contiguous()
        out = torch.zeros(1, dtype=torch.float32, device=a.device)
        da = torch.empty_like(af)
        check(profiler.launch("l1_mse", lambda: lib().pmoe_l1_mse(af.data_ptr(), bf.data_ptr(), af.numel(), int(is_mse), float(coef),
                                                                  out.data_ptr(), da.data_ptr(), stream_ptr())), "l1_mse")
        ctx.save_for_backward(da)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (da,) = ctx.saved_tensors
        return g * da, None, None, None


def punet_loss(actions, speed_pred, actions_gt, speed_gt, loss_coefs):
    """loss.py:135-142."""
    return _L1MseFn.apply(actions, actions_gt, False, loss_coefs[0]) + _L1MseFn.apply(speed_pred, speed_gt, True, loss_coefs[1])


def pmoe_loss(actions, speed_pred, actions_gt, speed_gt, loss_coefs):
    """loss.py:145-151 (the other arguments are dummies there too)."""
    return _L1MseFn.apply(actions, actions_gt, False, 1.0)


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, wce, wtv):
        _lib.require_cuda(pred, "pred")
        if pred.dtype != torch.float32:
            pred = pred.float()
        B, Cc, H, W = pred.shape
        tgt = target if target.dtype == torch.int64 else target.long()
        ws = torch.empty(lib().pmoe_segloss_workspace_floats(Cc, W), dtype=torch.float32, device=pred.device)
        out = torch.zeros(3, dtype=torch.float32, device=pred.device)
        check(profiler.launch("segloss_fwd", lambda: lib().pmoe_segloss_fwd(
            pred.data_ptr(), pred.stride(0), pred.stride(1), pred.stride(2), pred.stride(3), tgt.data_ptr(), tgt.stride(0),
            tgt.stride(1), tgt.stride(2), B, Cc, H, W, float(wce), float(wtv), ws.data_ptr(), out.data_ptr(), stream_ptr())),
            "segloss_fwd")
        ctx.save_for_backward(pred, tgt, ws)
        ctx.wce = float(wce)
        ctx.parts = out
        return out[0]

    @staticmethod
    def backward(ctx, g):
        pred, tgt, ws = ctx.saved_tensors
        B, Cc, H, W = pred.shape
        d = torch.empty_like(pred, memory_format=torch.contiguous_format)
        gs = g.detach().reshape(1).float().contiguous()
        check(profiler.launch("segloss_bwd", lambda: lib().pmoe_segloss_bwd(
            pred.data_ptr(), pred.stride(0), pred.stride(1), pred.stride(2), pred.stride(3), tgt.data_ptr(), tgt.stride(0),
            tgt.stride(1), tgt.stride(2), B, Cc, H, W, ctx.wce, ws.data_ptr(), gs.data_ptr(), 1.0, d.data_ptr(), d.stride(0),
            d.stride(1), d.stride(2), d.stride(3), 0, stream_ptr())), "segloss_bwd")
        return d, None, None, None


def cross_entropy_tversky_weighted_loss(pred, target, cross_entropy_weight=0.5, tversky_weight=0.5):
    """loss.py:47-55: dice-weighted cross entropy + Tversky, fused."""
    if cross_entropy_weight + tversky_weight != 1:
        raise ValueError("Cross Entropy weight and Tversky weight should " "sum to 1")
    return _SegLossFn.apply(pred, target, cross_entropy_weight, tversky_weight)


class AutoregressiveCriterion(nn.Module):
    """Per-frame sum of the segmentation loss with BPTT (loss.py:86-118). Only loss_type='tversky' (the reference
    default and the one stage 1 uses) runs on the fused kernels."""

    def __init__(self, n_target_frames: int = 1, loss_type: str = "tversky"):
        super().__init__()
        if loss_type != "tversky":
            raise NotImplementedError("pmoe_b200 AutoregressiveCriterion: only loss_type='tversky' is implemented")
        self.n_target_frames = n_target_frames
        self.loss_type = loss_type
        self.loss = cross_entropy_tversky_weighted_loss

    def forward(self, inputs, targets):
        assert inputs.size(1) == self.n_target_frames
        assert targets.size(1) == self.n_target_frames
        final_loss = 0
        for t in range(self.n_target_frames):
            final_loss = final_loss + self.loss(inputs[:, t, ...], targets[:, t, ...])
        return final_loss
