"""Host-side helpers the reference models import from utils.nn (reference: PMoE/utils/nn.py:10-58)."""
from typing import List

import torch
import torch.nn as nn


def freeze(model: nn.Module, exclude: List = [], verbose: bool = False) -> nn.Module:
    """Same semantics as the reference freeze(): with an empty exclusion list everything is frozen; otherwise every
    parameter whose NAME contains none of the `exclude` substrings is frozen (utils/nn.py:22-58)."""
    if exclude is None or len(exclude) == 0:
        for _, p in model.named_parameters():
            p.requires_grad_(False)
        if verbose:
            print(f"The whole model with {len(list(model.parameters()))} layers have been frozen.")
        return model
    frozen = []
    for name, p in model.named_parameters():
        if not any(layer in name for layer in exclude):
            frozen.append(name)
            p.requires_grad_(False)
    if verbose:
        print(f"{len(frozen)} layers have been frozen.")
    return model


def check_grad_norm(net: nn.Module) -> float:
    """Global L2 norm of all gradients (utils/nn.py:10-19) with ONE host sync instead of one per parameter."""
    from ..optim import check_grad_norm as _fused
    if any(p.grad is not None and p.grad.is_cuda for p in net.parameters()):
        return _fused(net)   # one multi-tensor launch (pmoe_mt_sqnorm)
    grads = [p.grad for p in net.parameters() if p.grad is not None]
    if not grads:
        return 0.0
    return float(torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.detach(), 2) for g in grads]), 2))
