"""Optimizer side of the training step (reference: trainer/train_2.py:157-165, utils/nn.py:10-19).

`clip_grad_norm_`, `check_grad_norm` and `FusedAdam` do what torch.nn.utils.clip_grad_norm_, the reference's
check_grad_norm and torch.optim.Adam(amsgrad=...) do, as THREE multi-tensor launches over a device-resident chunk
table instead of ~5 elementwise launches per parameter tensor and one host sync per parameter
(534 parameter tensors for the 6-expert mixture). The state layout (`exp_avg`, `exp_avg_sq`, `max_exp_avg_sq`, `step`)
is torch.optim.Adam's, so optimizer checkpoints interchange (train_2.py:87-105, 346-370).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, packs, profiler
from ._lib import check, lib, stream_ptr

_CHUNK = 1 << 14  # elements per chunk: 83 M parameters -> ~5 k chunks; one expert (13.8 M) still gives every SM several CTAs
_DT = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("vmax", "<u8"), ("n", "<i4"), ("pad", "<i4")])
assert _DT.itemsize == 48


class ChunkTable:
    """Device table of PmoeMtChunk records for a list of (param, grad, m, v, vmax) fp32 tensors; rebuilt only when a
    pointer changes (the caching allocator hands the same gradient addresses back in steady state)."""

    def __init__(self):
        self.key, self.dev, self.n = None, None, 0

    def get(self, rows, device):
        key = tuple((t.data_ptr() if t is not None else 0) for row in rows for t in row) + tuple(row[1].numel() for row in rows)
        if key == self.key:
            return self.dev, self.n
        recs = []
        for (p, g, m, v, vm) in rows:
            n = g.numel()
            for off in range(0, n, _CHUNK):
                b = off * 4
                recs.append(((p.data_ptr() + b) if p is not None else 0, g.data_ptr() + b, (m.data_ptr() + b) if m is not None else 0,
                             (v.data_ptr() + b) if v is not None else 0, (vm.data_ptr() + b) if vm is not None else 0,
                             min(_CHUNK, n - off), 0))
        arr = np.array(recs, dtype=_DT)
        host = torch.from_numpy(arr.view(np.uint8).copy())
        self.dev = host.to(device, non_blocking=False)
        self.n = len(recs)
        self.key = key
        return self.dev, self.n


def _flat_f32(t, what):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise RuntimeError("pmoe_b200.optim: %s must be a contiguous fp32 CUDA tensor" % what)
    _lib.require_cuda(t, what)
    return t


_norm_tables = {}


def grad_sqnorm(params, table=None):
    """fp64 device scalar: sum of squares of all gradients (one launch, no host sync)."""
    grads = [_flat_f32(p.grad, "gradient") for p in params if p.grad is not None]
    if not grads:
        return None, None, 0
    dev = grads[0].device
    table = table if table is not None else _norm_tables.setdefault(dev, ChunkTable())
    tdev, n = table.get([(None, g, None, None, None) for g in grads], dev)
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    check(profiler.launch("mt_sqnorm", lambda: lib().pmoe_mt_sqnorm(tdev.data_ptr(), n, out.data_ptr(), stream_ptr())), "mt_sqnorm")
    return out, tdev, n


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ (train_2.py:160-161) for the 2-norm: returns the total norm as a device
    scalar; the clip coefficient is computed and applied on the device."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("pmoe_b200 clip_grad_norm_: only the 2-norm the reference uses")
    params = [parameters] if isinstance(parameters, torch.Tensor) else list(parameters)
    sq, tdev, n = grad_sqnorm(params)
    if sq is None:
        return torch.zeros(())
    check(profiler.launch("mt_clip", lambda: lib().pmoe_mt_clip(tdev.data_ptr(), n, sq.data_ptr(), float(max_norm), stream_ptr())), "mt_clip")
    return sq.sqrt().float()[0]


def check_grad_norm(net):
    """Global L2 norm of all gradients (utils/nn.py:10-19): one kernel and ONE host sync."""
    sq, _, _ = grad_sqnorm(list(net.parameters()))
    return 0.0 if sq is None else float(sq.sqrt().item())


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay, amsgrad) with a single-launch step.
    step(max_grad_norm=...) additionally folds clip_grad_norm_ into the update (the gradients are read scaled, not
    rewritten)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sq = None
        if max_grad_norm is not None:
            sq, _, _ = grad_sqnorm([p for g in self.param_groups for p in g["params"]])
        for gi, group in enumerate(self.param_groups):
            by_step = {}   # torch.optim.Adam keeps one step counter per parameter: one launch per distinct value
            for p in group["params"]:
                if p.grad is None:
                    continue
                _flat_f32(p, "parameter")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["amsgrad"]:
                        st["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                step_no = int(st["step"].item())   # host tensor: no device sync
                by_step.setdefault(step_no, []).append((p, _flat_f32(p.grad, "gradient"), st["exp_avg"], st["exp_avg_sq"],
                                                        st.get("max_exp_avg_sq")))
            b1, b2 = group["betas"]
            for step_no, rows in by_step.items():
                dev = rows[0][0].device
                tdev, n = self._tables.setdefault((gi, dev, step_no if len(by_step) > 1 else -1), ChunkTable()).get(rows, dev)
                check(profiler.launch("mt_adam", lambda: lib().pmoe_mt_adam(
                    tdev.data_ptr(), n, float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                    step_no, int(group["amsgrad"]), None if sq is None else sq.data_ptr(), float(max_grad_norm or 0.0), stream_ptr())),
                    "mt_adam")
                packs.bump([r[0] for r in rows])   # the kernel wrote the parameters through raw pointers
        return loss


class FusedRMSprop(torch.optim.Optimizer):
    """torch.optim.RMSprop(params, lr, alpha, eps, weight_decay, momentum, centered) — the reference's alternative
    optimizer (train_2.py:67-71, conf/stage_2.yaml:147-153: centered, alpha 0.99) — with a single-launch step. State
    names (`square_avg`, `grad_avg`, `momentum_buffer`, `step`) are torch's, so optimizer checkpoints interchange.
    step(max_grad_norm=...) folds clip_grad_norm_ into the update like FusedAdam."""

    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0.0, momentum=0.0, centered=False):
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay, momentum=momentum, centered=centered))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sq = None
        if max_grad_norm is not None:
            sq, _, _ = grad_sqnorm([p for g in self.param_groups for p in g["params"]])
        for gi, group in enumerate(self.param_groups):
            rows = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                _flat_f32(p, "parameter")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["momentum"] > 0:
                        st["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if group["centered"]:
                        st["grad_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                rows.append((p, _flat_f32(p.grad, "gradient"), st["square_avg"], st.get("momentum_buffer"), st.get("grad_avg")))
            if not rows:
                continue
            dev = rows[0][0].device
            tdev, n = self._tables.setdefault((gi, dev), ChunkTable()).get(rows, dev)
            check(profiler.launch("mt_rmsprop", lambda: lib().pmoe_mt_rmsprop(
                tdev.data_ptr(), n, float(group["lr"]), float(group["alpha"]), float(group["eps"]), float(group["weight_decay"]),
                float(group["momentum"]), None if sq is None else sq.data_ptr(), float(max_grad_norm or 0.0), stream_ptr())),
                "mt_rmsprop")
            packs.bump([r[0] for r in rows])   # the kernel wrote the parameters through raw pointers
        return loss


class AveragedModel(torch.optim.swa_utils.AveragedModel):
    """torch.optim.swa_utils.AveragedModel (train_2.py:119-121) whose update_parameters (train_2.py:179-187) runs as ONE
    multi-tensor launch over all parameters instead of one lerp per parameter tensor. Default averaging only
    (equal-weight running mean, parameters only — what the reference constructs); anything else defers to torch."""

    def __init__(self, model, device=None, avg_fn=None, multi_avg_fn=None, use_buffers=False):
        super().__init__(model, device, avg_fn, multi_avg_fn, use_buffers)
        self._fused = avg_fn is None and multi_avg_fn is None and not use_buffers
        self._table = ChunkTable()

    @torch.no_grad()
    def update_parameters(self, model):
        mine, theirs = list(self.module.parameters()), list(model.parameters())
        ok = self._fused and all(a.is_cuda and b.is_cuda and a.dtype == torch.float32 and b.dtype == torch.float32
                                 and a.is_contiguous() and b.is_contiguous() for a, b in zip(mine, theirs))
        if not ok:
            return super().update_parameters(model)
        n_avg = int(self.n_averaged.item())   # once per epoch
        if n_avg == 0:
            torch._foreach_copy_(mine, theirs)
        else:
            dev = mine[0].device
            tdev, n = self._table.get([(a, b.detach(), None, None, None) for a, b in zip(mine, theirs)], dev)
            check(profiler.launch("mt_swa_update", lambda: lib().pmoe_mt_swa_update(tdev.data_ptr(), n, n_avg, stream_ptr())),
                  "mt_swa_update")
            packs.bump(mine)
        # use_buffers=False: torch keeps the averaged module's buffers (BatchNorm running statistics,
        # num_batches_tracked) in sync with the model's on every update (swa_utils.py: `b_swa.copy_(b_model)`)
        b_mine, b_theirs = list(self.module.buffers()), list(model.buffers())
        if b_mine:
            torch._foreach_copy_(b_mine, [b.detach().to(a.device) for a, b in zip(b_mine, b_theirs)])
        self.n_averaged += 1
