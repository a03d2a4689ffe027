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


class _OneHotLossFn(torch.autograd.Function):
    """L1 / MSE / L1+gradient-difference of fp32 NCHW logits against the one-hot of an int64 class map (onehot_loss.cu)."""

    @staticmethod
    def forward(ctx, pred, target, mode):
        _lib.require_cuda(pred, "pred")
        if pred.dtype != torch.float32:
            pred = pred.float()
        B, Cc, H, W = pred.shape
        tgt = target if target.dtype == torch.int64 else target.long()
        sums = torch.empty(2, dtype=torch.float64, device=pred.device)
        out = torch.empty(1, dtype=torch.float32, device=pred.device)
        check(profiler.launch("onehot_loss_fwd", lambda: lib().pmoe_onehot_loss_fwd(
            pred.data_ptr(), pred.stride(0), pred.stride(1), pred.stride(2), pred.stride(3), tgt.data_ptr(), tgt.stride(0),
            tgt.stride(1), tgt.stride(2), B, Cc, H, W, int(mode), sums.data_ptr(), out.data_ptr(), stream_ptr())),
            "onehot_loss_fwd")
        ctx.save_for_backward(pred, tgt)
        ctx.mode = int(mode)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        pred, tgt = ctx.saved_tensors
        B, Cc, H, W = pred.shape
        d = torch.empty_like(pred, memory_format=torch.contiguous_format)
        gs = g.detach().reshape(1).float().contiguous()
        check(profiler.launch("onehot_loss_bwd", lambda: lib().pmoe_onehot_loss_bwd(
            pred.data_ptr(), pred.stride(0), pred.stride(1), pred.stride(2), pred.stride(3), tgt.data_ptr(), tgt.stride(0),
            tgt.stride(1), tgt.stride(2), B, Cc, H, W, ctx.mode, gs.data_ptr(), d.data_ptr(), d.stride(0), d.stride(1),
            d.stride(2), d.stride(3), stream_ptr())), "onehot_loss_bwd")
        return d, None, None


def l1_gdl(inputs: torch.Tensor, targets: torch.Tensor):
    """loss.py:58-83: L1 + gradient-difference loss of the LAST frame's raw logits against the one-hot target.
    inputs (B,T,C,H,W) fp32, targets (B,T,H,W) int64."""
    return _OneHotLossFn.apply(inputs[:, -1], targets[:, -1], 2)


def _class_counts(pred, target):
    """The per-class dice weights 1 - 2(|P∩T|+eps)/(|P|+|T|+eps) of loss.py:6-17 from the fused forward pass."""
    _lib.require_cuda(pred, "pred")
    if pred.dtype != torch.float32:
        pred = pred.float()
    B, Cc, H, W = pred.shape
    tgt = target if target.dtype == torch.int64 else target.long()
    ws = torch.empty(lib().pmoe_segloss_workspace_floats(Cc, W), dtype=torch.float32, device=pred.device)
    out = torch.zeros(3, dtype=torch.float32, device=pred.device)
    check(profiler.launch("segloss_fwd", lambda: lib().pmoe_segloss_fwd(
        pred.data_ptr(), pred.stride(0), pred.stride(1), pred.stride(2), pred.stride(3), tgt.data_ptr(), tgt.stride(0),
        tgt.stride(1), tgt.stride(2), B, Cc, H, W, 0.5, 0.5, ws.data_ptr(), out.data_ptr(), stream_ptr())), "segloss_fwd")
    return ws[4 * Cc:5 * Cc]


@torch.no_grad()
def class_dice(pred, target, epsilon=1e-6):
    """loss.py:6-17: 1 - dice per class of the argmax prediction (the CE weights), one pass instead of a 23-iteration loop."""
    if epsilon != 1e-6:
        raise NotImplementedError("pmoe_b200 class_dice: epsilon is fixed at the reference default 1e-6")
    return _class_counts(pred.detach(), target).clone()


@torch.no_grad()
def dice_score(pred, target, epsilon=1e-6):
    """loss.py:20-31: per-class dice of the argmax prediction — the validation metric of train_0.py:230 / train_1.py:249."""
    if epsilon != 1e-6:
        raise NotImplementedError("pmoe_b200 dice_score: epsilon is fixed at the reference default 1e-6")
    return 1.0 - _class_counts(pred.detach(), target)


def tversky_loss(pred, target, alpha=0.5, beta=0.5):
    """loss.py:34-44 on its own: the Tversky half of the fused kernel (CE weight 0)."""
    if alpha != 0.5 or beta != 0.5:
        raise NotImplementedError("pmoe_b200 tversky_loss: alpha = beta = 0.5 (the only values the reference uses)")
    return _SegLossFn.apply(pred, target, 0.0, 1.0)


class AutoregressiveCriterion(nn.Module):
    """Per-frame sum of the per-frame loss with BPTT (loss.py:86-118): 'tversky' (reference default, the fused dice-weighted
    CE + Tversky kernels), 'l1' / 'l2' (nn.L1Loss / nn.MSELoss against the one-hot target, fused: no one-hot tensor)."""

    def __init__(self, n_target_frames: int = 1, loss_type: str = "tversky"):
        super().__init__()
        self.n_target_frames = n_target_frames
        self.loss_type = loss_type
        if loss_type == "l1":
            self.loss = lambda x, t: _OneHotLossFn.apply(x, t, 0)
        elif loss_type == "l2":
            self.loss = lambda x, t: _OneHotLossFn.apply(x, t, 1)
        elif loss_type == "tversky":
            self.loss = cross_entropy_tversky_weighted_loss
        else:
            raise ValueError(f"Unknown loss type {loss_type}, supported ones are L1, L2, and tversky")

    def forward(self, inputs, targets):
        assert inputs.size(1) == self.n_target_frames
        assert targets.size(1) == self.n_target_frames
        final_loss = 0
        for t in range(self.n_target_frames):
            final_loss = final_loss + self.loss(inputs[:, t, ...], targets[:, t, ...])
        return final_loss
