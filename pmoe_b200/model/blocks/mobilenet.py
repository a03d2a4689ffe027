"""MobileNetV2 / V3-Small / V3-Large backbones with the reference's ECA stem (reference: PMoE/model/blocks/backbone.py:75-104).

Like the reference, the module tree IS torchvision's (so state_dict keys, BatchNorm eps / momentum and the block configuration
are torchvision's by construction): `features[0][0]` is replaced by a stride-1 EfficientConvBlock and the classifier's last
Linear by one with 512 outputs. torchvision only holds the parameters; `forward` walks that tree and records every layer on the
pmoe_b200 tape — 1x1 convolutions on the tensor-core kernel, depthwise k x k convolutions / squeeze-excite / ReLU6 / Hardswish /
Hardsigmoid on the kernels of csrc/depthwise.cu, eltwise*.cu — so forward AND backward run on the library, as for the ResNets.
"""
import torch
import torch.nn as nn

from ... import nhwc, train
from .basics import EfficientConvBlock

_ACT = {nn.ReLU: "relu", nn.ReLU6: "relu6", nn.Hardswish: "hswish", nn.Hardsigmoid: "hsigmoid"}


def _act_of(m):
    for cls, name in _ACT.items():
        if isinstance(m, cls):
            return name
    return None


def _units(seq):
    """Flatten an InvertedResidual's Sequential into units: ("conv", Conv2d, BatchNorm2d | None, act | None) or ("se", module)."""
    out, mods, i = [], list(seq), 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Sequential) and len(m) and isinstance(m[0], nn.Conv2d):   # Conv2dNormActivation
            sub = list(m)
            bn = sub[1] if len(sub) > 1 and isinstance(sub[1], nn.BatchNorm2d) else None
            act = _act_of(sub[-1]) if len(sub) > (2 if bn is not None else 1) else None
            out.append(("conv", sub[0], bn, act))
            i += 1
        elif isinstance(m, nn.Conv2d):                                                # bare conv (+ BatchNorm2d): MobileNetV2's projection
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d) else None
            out.append(("conv", m, bn, None))
            i += 2 if bn is not None else 1
        elif hasattr(m, "fc1") and hasattr(m, "fc2"):                                 # torchvision.ops.SqueezeExcitation
            if not isinstance(m.activation, nn.ReLU) or not isinstance(m.scale_activation, nn.Hardsigmoid):
                raise NotImplementedError("pmoe_b200 mobilenet: squeeze-excite with ReLU / Hardsigmoid only")
            out.append(("se", m))
            i += 1
        else:
            raise NotImplementedError("pmoe_b200 mobilenet: unexpected layer %r" % (m,))
    return out


def _conv_unit(tape, conv, bn, act, x, residual, want_pool, tag):
    if conv.groups == 1:
        if conv.kernel_size != (1, 1) or conv.stride != (1, 1):
            raise NotImplementedError("pmoe_b200 mobilenet: dense convolutions inside the blocks are 1x1 / stride 1 (got %r)" % (conv,))
        return train.conv_op(tape, [x], conv.weight, conv.bias, bn, act, residual=residual, want_pool=want_pool, ksize=1, tag=tag)
    if residual is not None or want_pool:
        raise NotImplementedError("pmoe_b200 mobilenet: a depthwise convolution never carries the residual add")
    return train.dwconv_op(tape, x, conv, bn, act, tag=tag), None


def inverted_residual(tape, blk, x, tag=""):
    """torchvision InvertedResidual (V2: `.conv`, V3: `.block`): [1x1 expand] -> depthwise -> [squeeze-excite] -> 1x1 project
    (+ x when `use_res_connect`); the add rides the projection's BatchNorm pass."""
    units = _units(blk.conv if hasattr(blk, "conv") else blk.block)
    y = x
    for j, u in enumerate(units):
        last = j == len(units) - 1
        if u[0] == "se":
            y = train.se_op(tape, u[1], y, tag="%s.%d" % (tag, j))
        else:
            y, _ = _conv_unit(tape, u[1], u[2], u[3], y, x if (last and blk.use_res_connect) else None, False, "%s.%d" % (tag, j))
    return y


class MobileNetECA(nn.Module):
    """torchvision MobileNetV2 / V3 with `features[0][0] := EfficientConvBlock(stride 1)` and a 512-wide last Linear."""

    def __init__(self, arch, in_ch, gamma=2, b=1, pretrained=False):
        super().__init__()
        import torchvision
        ctor = {"mobilenet_v3_small": torchvision.models.mobilenet_v3_small, "mobilenet_v3_large": torchvision.models.mobilenet_v3_large,
                "mobilenet_v2": torchvision.models.mobilenet_v2}.get(arch, torchvision.models.mobilenet_v3_small)   # backbone.py:86-90
        try:
            net = ctor(weights="IMAGENET1K_V1" if pretrained else None)
        except Exception as ex:
            raise RuntimeError("pmoe_b200 get_backbone(pretrained=True): torchvision could not provide the ImageNet weights of %s (%s)"
                               % (arch, str(ex).splitlines()[0][:200])) from ex
        net.features[0][0] = EfficientConvBlock(in_ch=in_ch, out_ch=net.features[0][0].out_channels, gamma=gamma, b=b)
        if "v2" in arch:
            net.classifier = nn.Linear(net.classifier[1].in_features, 512)           # backbone.py:98-99
        else:
            net.classifier[3] = nn.Linear(net.classifier[3].in_features, 512)        # backbone.py:100-101
        self.features, self.classifier = net.features, net.classifier   # the global average pool has no parameters

    def tape_features(self, tape, x, tag="backbone", layout=None, pool_in=None):
        """NHWC input Act -> (1,1,B,512) feature Act. layout / pool_in: channel-group layout and per-image channel sums of x when it
        is a window of the PU-Net mask ring (PUNetExpert, moe.py:303-317)."""
        if x.t.shape[1] % 32 or x.t.shape[2] % 32:
            raise RuntimeError("pmoe_b200 MobileNet backbone needs H and W divisible by 32 (got %dx%d)" % (x.t.shape[1], x.t.shape[2]))
        stem = list(self.features[0])
        bn0, act0 = stem[1], _act_of(stem[2])
        y = train.eca_conv_block(tape, stem[0], x, layout, pool_in, tag=tag + ".features.0.0", want_out_stats=bn0.training)
        y = train.bn_act_op(tape, bn0, y, act0, tag=tag + ".features.0.1")
        n_feat = len(self.features)
        for i in range(1, n_feat - 1):
            y = inverted_residual(tape, self.features[i], y, tag="%s.features.%d" % (tag, i))
        (_, conv, bn, act), = _units([self.features[n_feat - 1]])
        y, pool = _conv_unit(tape, conv, bn, act, y, None, True, "%s.features.%d" % (tag, n_feat - 1))
        feat = train.feature_act(tape, train.InterRepr(tape, y, pool))               # adaptive_avg_pool2d(1) + flatten
        mods = [self.classifier] if isinstance(self.classifier, nn.Linear) else list(self.classifier)
        for j, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                feat = train.linear_op(tape, [feat], m, None, tag="%s.classifier.%d" % (tag, j))
            elif isinstance(m, nn.Dropout):
                feat = train.dropout_op(tape, feat, m.p, m.training)
            elif _act_of(m) is not None:
                feat = train.act_op(tape, feat, _act_of(m), tag="%s.classifier.%d" % (tag, j))
            else:
                raise NotImplementedError("pmoe_b200 mobilenet: unexpected classifier layer %r" % (m,))
        return feat

    def forward(self, x):
        """x: fp32 (B, C, H, W) on the GPU -> (B, 512) features."""
        def runner(tape):
            a = nhwc.from_nchw(x, dtype=tape.dtype)
            feat = self.tape_features(tape, a)

            def seed(tp, g):
                if g[0] is not None:
                    train.seed_vec(tp, feat, g[0])
            return [train.vec_value(feat)], seed
        return train.run(self, runner)[0]
