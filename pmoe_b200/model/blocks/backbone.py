"""Backbone factory with the reference's signature (reference: PMoE/model/blocks/backbone.py:13-72).

`resnet18` is the conf default (conf/stage_2.yaml:111) and the only architecture implemented on the B200
kernels so far; the module reproduces torchvision's ResNet-18 state_dict keys with `conv1` replaced by an
EfficientConvBlock and `fc` by Identity, so reference checkpoints load with strict=True.
"""
import torch
import torch.nn as nn

from ... import nhwc, train
from .basics import EfficientConvBlock


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        self.stride = stride


class ResNet18ECA(nn.Module):
    def __init__(self, in_ch, gamma=2, b=1):
        super().__init__()
        self.conv1 = EfficientConvBlock(in_ch=in_ch, out_ch=64, gamma=gamma, b=b)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        planes = 64
        for i, (c, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            setattr(self, "layer%d" % i, nn.Sequential(BasicBlock(planes, c, stride), BasicBlock(c, c, 1)))
            planes = c
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Identity()

    def forward(self, x):
        """x: fp32 (B, C, H, W) on the GPU -> (B, 512) features."""
        def runner(tape):
            a = nhwc.from_nchw(x, dtype=tape.dtype)
            feat = train.resnet18_eca(tape, self, a)
            return [feat.value], (lambda tp, g: feat.backward(g[0]) if g[0] is not None else None)
        return train.run(self, runner)[0]


def get_backbone(arch: str = "resnet18", n_frames: int = 4, pretrained: bool = False, gamma: int = 2, b: int = 1,
                 n_channels: int = 3):
    if arch.lower() != "resnet18":
        raise NotImplementedError("pmoe_b200 get_backbone: only 'resnet18' (the conf default) runs on the B200 kernels; got %r" % arch)
    if pretrained:
        raise RuntimeError("pmoe_b200 get_backbone: pretrained=True needs the ImageNet download (no network here); load a "
                           "checkpoint with load_state_dict instead and pass pretrained=False")
    return ResNet18ECA(n_frames * n_channels, gamma, b)


def get_unet(*args, **kwargs):
    raise NotImplementedError("pmoe_b200: the 'segmentation' backbone type is broken in the reference itself "
                              "(model/moe.py:95 concatenates the tuple returned by get_unet(inter_repr=True))")
