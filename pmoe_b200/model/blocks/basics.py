"""Building blocks with the reference's constructor signatures and state_dict keys
(reference: PMoE/model/blocks/basics.py). The nn.Conv2d / nn.BatchNorm2d / nn.Linear children are
parameter containers only — forward() never calls them; it runs the sm_100a kernels.
"""
from collections import OrderedDict
from math import log2
from typing import List

import torch
import torch.nn as nn

from ... import config, infer, nhwc, train

_ACT_NAMES = {"relu": "relu", "tanh": "tanh", "sigmoid": "sigmoid", "elu": "elu"}


def _grad_mode(module):
    """True -> run the general (train-capable) path: BatchNorm in batch-statistics mode and/or autograd.
    False -> the fused eval path (BN folded into the conv epilogue, nothing saved for backward)."""
    if module.training:
        return True
    return torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters())


class MLP(nn.Sequential):
    """Result of make_mlp: same child indices as the reference's nn.Sequential (basics.py:11-45)."""

    def forward(self, x):
        return train.mlp_forward(self, x)


def make_mlp(dims: List, act: str, l_act: bool = False, bn: bool = True, dropout: float = 0.0):
    """Same arguments and child layout as reference make_mlp (basics.py:11-45)."""
    act_mod = {"relu": nn.ReLU(inplace=True), "tanh": nn.Tanh(), "sigmoid": nn.Sigmoid(), "elu": nn.ELU()}[act.lower()]
    children = []
    last = len(dims) - 2
    for i, (d_in, d_out) in enumerate(zip(dims[:-1], dims[1:])):
        children.append(nn.Linear(d_in, d_out, bias=not bn))
        if i != last:
            if bn:
                children.append(nn.BatchNorm1d(d_out))
            children.append(act_mod)
            if dropout > 0.0:
                children.append(nn.Dropout(p=dropout))
    if l_act:
        children.append(act_mod)
    seq = MLP(*children)
    seq.act_name = _ACT_NAMES[act.lower()]
    return seq


class Conv3Block(nn.Sequential):
    """conv3 (basics.py:48-59): children 0,1,3,4 carry the parameters; 2,5 are the ReLU placeholders."""

    def forward(self, x):  # x: NCHW fp32 (stand-alone use); the enclosing networks call the NHWC paths directly
        if _grad_mode(self):
            return train.nhwc_module_forward(self, x, lambda tape, a: train.conv3_block(tape, self, [a])[0])
        y = infer.conv3_block_eval(self, [nhwc.from_nchw(x, dtype=config.act_dtype())])
        return nhwc.to_nchw(y.t, y.c)


def conv3(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    if stride != 1:
        raise NotImplementedError("pmoe_b200 conv3: the reference only ever uses stride 1 (basics.py:48)")
    return Conv3Block(
        nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1, bias=False), nn.BatchNorm2d(out_ch),
        nn.ReLU(inplace=True),
        nn.Conv2d(out_ch, out_ch, kernel_size=3, stride=stride, padding=1, bias=False), nn.BatchNorm2d(out_ch),
        nn.ReLU(inplace=True))


def eca_kernel_size(channels: int, gamma: int = 2, b: int = 1) -> int:
    t = int(abs((log2(channels) + b) / gamma))
    return t if t % 2 else t + 1


class EfficientBlock(nn.Module):
    """ECA channel attention (basics.py:62-77); `conv` holds the (1,1,k) kernel."""

    def __init__(self, channels: int, gamma: int = 2, b: int = 1):
        super().__init__()
        k = eca_kernel_size(channels, gamma, b)
        self.conv = nn.Conv1d(1, 1, kernel_size=k, padding=int(k / 2), bias=False)

    def forward(self, x):
        if _grad_mode(self):
            return train.nhwc_module_forward(self, x, lambda tape, a: train.eca_op(tape, self, a))
        a = nhwc.from_nchw(x, dtype=config.act_dtype())
        sums = nhwc.channel_sums(a.t)
        gate = nhwc.eca_gate(sums, a.t.shape[1] * a.t.shape[2], self.conv.weight, 1, a.c, a.cpad)
        return nhwc.to_nchw(nhwc.scale_channels(a.t, gate), a.c)


class EfficientConvBlock(nn.Module):
    """Two-layer ECA conv stem (basics.py:80-135): layer1.{eca1,conv1.{0,1}}, layer2.{eca2,conv2.{0,1}}."""

    def __init__(self, in_ch: int, out_ch: int, stride: int = 1, gamma: int = 2, b: int = 1):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("pmoe_b200 EfficientConvBlock: stride is always 1 in the reference's use")
        self.in_ch, self.out_ch = in_ch, out_ch
        self.layer1 = nn.Sequential(OrderedDict([
            ("eca1", EfficientBlock(in_ch, gamma, b)),
            ("conv1", nn.Sequential(nn.Conv2d(in_ch, 64, kernel_size=3, stride=1, padding=1, bias=False),
                                    nn.BatchNorm2d(64), nn.ReLU(inplace=True)))]))
        self.layer2 = nn.Sequential(OrderedDict([
            ("eca2", EfficientBlock(64, gamma, b)),
            ("conv2", nn.Sequential(nn.Conv2d(64, out_ch, kernel_size=3, stride=1, padding=1, bias=False),
                                    nn.BatchNorm2d(out_ch), nn.ReLU(inplace=True)))]))

    def forward(self, x):
        if _grad_mode(self):
            return train.nhwc_module_forward(self, x, lambda tape, a: train.eca_conv_block(tape, self, a))
        a = nhwc.from_nchw(x, dtype=config.act_dtype())
        sums = nhwc.channel_sums(a.t)
        y = infer.eca_conv_block_eval(self, a.t, (1, a.c, a.cpad), sums, a.t.shape[1] * a.t.shape[2])
        return nhwc.to_nchw(y.t, y.c)
