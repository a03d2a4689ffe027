"""Backbone factory with the reference's signature (reference: PMoE/model/blocks/backbone.py:13-72).

`resnet18` is the conf default (conf/stage_2.yaml:111); `resnet34` (BasicBlock, [3,4,6,3]) and `resnet50` (Bottleneck,
[3,4,6,3], fc := Linear(2048, 512)) are the other ResNets `_get_resnet` accepts. The modules reproduce torchvision's
state_dict keys with `conv1` replaced by an EfficientConvBlock and `fc` by Identity / Linear, so reference checkpoints load
with strict=True. `get_unet` builds the 'segmentation' backbone (entry block + U-Net). The mobilenet family (`_get_mobilenet`,
backbone.py:75-104: depthwise convolutions, ReLU6 / Hardswish / Hardsigmoid, squeeze-excite) lives in `mobilenet.py`.
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


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)   # torchvision v1.5: stride on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or inplanes != planes * 4:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride, bias=False), nn.BatchNorm2d(planes * 4))
        self.stride = stride


_ARCH = {"resnet18": (BasicBlock, (2, 2, 2, 2)), "resnet34": (BasicBlock, (3, 4, 6, 3)), "resnet50": (Bottleneck, (3, 4, 6, 3))}


class ResNet18ECA(nn.Module):
    """ResNet-18/34/50 with the EfficientConvBlock stem (the class keeps its first name; `arch` selects the stages)."""

    def __init__(self, in_ch, gamma=2, b=1, arch="resnet18"):
        super().__init__()
        block, counts = _ARCH[arch]
        self.conv1 = EfficientConvBlock(in_ch=in_ch, out_ch=64, gamma=gamma, b=b)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        planes = 64
        for i, ((c, stride), nb) in enumerate(zip(((64, 1), (128, 2), (256, 2), (512, 2)), counts), start=1):
            blocks = [block(planes, c, stride)]
            planes = c * block.expansion
            blocks += [block(planes, c, 1) for _ in range(nb - 1)]
            setattr(self, "layer%d" % i, nn.Sequential(*blocks))
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Identity() if planes == 512 else nn.Linear(planes, 512)   # backbone.py:66-69

    def forward(self, x):
        """x: fp32 (B, C, H, W) on the GPU -> (B, 512) features."""
        def runner(tape):
            a = nhwc.from_nchw(x, dtype=tape.dtype)
            feat = train.backbone_features(tape, self, a)

            def seed(tp, g):
                if g[0] is not None:
                    train.seed_vec(tp, feat, g[0])
            return [train.vec_value(feat)], seed
        return train.run(self, runner)[0]


def get_backbone(arch: str = "resnet18", n_frames: int = 4, pretrained: bool = False, gamma: int = 2, b: int = 1,
                 n_channels: int = 3):
    if "resnet" in arch:           # backbone.py:21-25
        key = arch.lower() if arch.lower() in _ARCH else "resnet18"   # backbone.py:57-61: unknown ResNet names fall back to resnet18
    elif "mobilenet" in arch:
        from .mobilenet import MobileNetECA
        return MobileNetECA(arch.lower(), n_frames * n_channels, gamma, b, pretrained)
    else:
        return None                # the reference's factory falls through (and its callers then fail on None)
    arch = key
    net = ResNet18ECA(n_frames * n_channels, gamma, b, arch.lower())
    if pretrained:
        _load_imagenet_weights(net, arch.lower())
    return net


def _load_imagenet_weights(net, arch):
    """`models.resnetXX(pretrained=True)` of the reference (backbone.py:57-61): torchvision's ImageNet weights for every layer
    the factory keeps — everything except `conv1` (replaced by the EfficientConvBlock) and `fc` (Identity / a fresh Linear).
    torchvision fetches the file into the torch hub cache on first use; without it (and without a network) this raises."""
    import torchvision
    try:
        ref = getattr(torchvision.models, arch)(weights="IMAGENET1K_V1")   # what the legacy `pretrained=True` selects
    except Exception as ex:  # no cached weights and no network
        raise RuntimeError("pmoe_b200 get_backbone(pretrained=True): torchvision could not provide the ImageNet weights of %s (%s: %s); "
                           "place the torchvision checkpoint in the torch hub cache or pass pretrained=False and load a checkpoint"
                           % (arch, type(ex).__name__, str(ex).splitlines()[0][:200])) from ex
    sd = {k: v for k, v in ref.state_dict().items() if not (k.startswith("conv1.") or k.startswith("fc."))}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    bad = [k for k in missing if not (k.startswith("conv1.") or k.startswith("fc."))]
    if bad or unexpected:
        raise RuntimeError("pmoe_b200 get_backbone(pretrained=True): key mismatch with torchvision's %s: %r %r" % (arch, bad, unexpected))


def get_unet(model_dir: str, inter_repr: bool = True, n_frames: int = 4, gamma: int = 2, b: int = 1, n_channels: int = 3):
    """backbone.py:28-45: EfficientConvBlock(n_frames*n_channels -> 3) followed by a U-Net whose weights come from
    `model_dir` (strict=False, exactly as the reference loads them). NB with inter_repr=True the Sequential returns the
    tuple (bottleneck, logits); the reference's own experts then fail at `torch.cat` (model/moe.py:95) — that caller-side
    defect is not papered over here, the factory itself behaves like the reference's."""
    from pathlib import Path
    from .unet import UNet
    model = UNet(gamma=gamma, b=b, inter_repr=inter_repr)
    state_dict = torch.load(Path(model_dir).resolve(), map_location="cpu")
    model.load_state_dict(state_dict, strict=False)
    entry_block = EfficientConvBlock(in_ch=n_frames * n_channels, out_ch=3, gamma=gamma, b=b)
    return nn.Sequential(entry_block, model)
