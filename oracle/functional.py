"""TEST INFRASTRUCTURE — CPU oracle for the pmoe_b200 hot path. NOT product code.

A functional (state-dict driven) fp32 PyTorch restatement of the reference algorithm. Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package; pmoe_b200/ never does. Every function cites the reference lines it restates
(paths relative to /root/reference/PMoE). The restatement is pinned against outputs of the live
reference (tests/golden/*.pt, produced by oracle/gen_golden.py) in tests/test_oracle_golden.py.

Conventions: `sd` is a flat dict name -> tensor with the reference's state_dict keys, `p` a key
prefix ending in '.', `train` selects BatchNorm batch statistics (running stats in `sd` are then
updated in place exactly as nn.BatchNorm2d does).
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# ------------------------------------------------------------------ optional storage-precision emulation
# The product stores every activation AND every activation gradient in bf16 (fp32 accumulation inside the kernels). A ReLU
# network's gradient is a discontinuous function of its pre-activations, so comparing a bf16-storage step with the fp32
# restatement measures how many ReLU masks the storage rounding flips, not whether the kernels are right. With
# `storage("bf16")` active the SAME restatement rounds each stored tensor (forward value and, on the way back, its gradient)
# to bf16 at the points where the product stores one: conv / linear outputs, BatchNorm+activation(+residual) outputs, ECA
# scaled outputs, pooled features. Arithmetic stays fp32. Default (None): the plain fp32 restatement the goldens pin.
import contextlib

STORE = None


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


@contextlib.contextmanager
def storage(kind):
    global STORE
    old = STORE
    STORE = {"bf16": _RoundBf16.apply, None: None, "fp32": None}[kind]
    try:
        yield
    finally:
        STORE = old


def _st(x):
    return x if STORE is None else STORE(x)


class _BnStored(torch.autograd.Function):
    """conv output -> train-mode BatchNorm under bf16 storage, forward and backward as the product's kernels compute them: batch
    statistics of the UNROUNDED fp32 accumulators (conv epilogue), normalisation of the STORED (rounded) tensor r with them;
    backward by the standard formula on xhat = (r - mean) * rstd, the resulting gradient of the conv output stored in bf16."""

    @staticmethod
    def forward(ctx, src, weight, bias, eps):
        dims = [d for d in range(src.dim()) if d != 1]
        shape = [1, -1] + [1] * (src.dim() - 2)
        mean, var = src.mean(dim=dims), src.var(dim=dims, unbiased=False)
        rstd = torch.rsqrt(var + eps)
        r = src.to(torch.bfloat16).to(src.dtype)
        xhat = (r - mean.view(shape)) * rstd.view(shape)
        ctx.save_for_backward(xhat, weight, rstd)
        ctx.dims, ctx.shape = dims, shape
        return xhat * weight.view(shape) + bias.view(shape)

    @staticmethod
    def backward(ctx, g):
        xhat, weight, rstd = ctx.saved_tensors
        dims, shape = ctx.dims, ctx.shape
        s1, s2 = g.sum(dim=dims), (g * xhat).sum(dim=dims)
        n = g.numel() / g.shape[1]
        dx = (weight * rstd).view(shape) * (g - (s1 / n).view(shape) - xhat * (s2 / n).view(shape))
        return dx.to(torch.bfloat16).to(g.dtype), s2, s1, None


def _wq(w):
    """A conv / linear weight as the product's kernels read it under bf16 storage: rounded to bf16 (the packed operand)."""
    return w if STORE is None else w.to(torch.bfloat16).to(w.dtype)


def _conv(x, w, bias, stride, pad):
    """Dense convolution whose output the product stores: weights as stored (bf16 under storage emulation), output through
    `_stc`; the arguments ride along for the eval-mode BatchNorm fold (see `batchnorm`)."""
    if STORE is None:
        return F.conv2d(x, w, bias, stride, pad)
    y = _stc(F.conv2d(x, _wq(w), bias, stride, pad))
    y._pmoe_conv = (x, w, bias, stride, pad)
    return y


def _stc(x):
    """A stored convolution output. The product takes the train-mode BatchNorm statistics of a conv output from the fp32
    accumulators (conv epilogue) and normalises the STORED, rounded tensor with them; the rounded tensor therefore carries its
    unrounded source along for `batchnorm` (a 1.5e-3 difference in the normalised output otherwise)."""
    if STORE is None:
        return x
    y = STORE(x)
    y._pmoe_unrounded = x
    return y


# ------------------------------------------------------------------ blocks (model/blocks/basics.py)
def batchnorm(x, sd, p, train, eps=BN_EPS, momentum=BN_MOMENTUM):
    """nn.BatchNorm{1,2}d defaults as instantiated at basics.py:34,52,55,103,124 (torchvision's MobileNetV3 passes its own
    eps / momentum)."""
    if train and p + "num_batches_tracked" in sd:
        sd[p + "num_batches_tracked"] += 1
    src = getattr(x, "_pmoe_unrounded", None)
    if train and STORE is not None and src is not None:
        # storage emulation of conv -> train-mode BatchNorm exactly as the product computes it (see _BnStored)
        with torch.no_grad():
            dims = [d for d in range(src.dim()) if d != 1]
            n = src.numel() / src.shape[1]
            sd[p + "running_mean"].mul_(1 - momentum).add_(momentum * src.mean(dim=dims))
            sd[p + "running_var"].mul_(1 - momentum).add_(momentum * src.var(dim=dims, unbiased=False) * n / max(n - 1, 1))
        return _BnStored.apply(src, sd[p + "weight"], sd[p + "bias"], eps)
    if (not train) and STORE is not None and src is not None:
        # eval mode: the product folds the BatchNorm into the conv — scale into the packed bf16 weights, shift into the epilogue —
        # so the conv output is never stored on its own; the first rounding happens after BatchNorm + activation (the caller's _st)
        conv_args = getattr(x, "_pmoe_conv", None)
        if conv_args is not None and conv_args[2] is None:
            xin, w, _, stride, pad = conv_args
            scale = sd[p + "weight"] * torch.rsqrt(sd[p + "running_var"] + eps)
            shift = sd[p + "bias"] - sd[p + "running_mean"] * scale
            return F.conv2d(xin, _wq(w * scale.view(-1, 1, 1, 1)), None, stride, pad) + shift.view(1, -1, 1, 1)
        x = src
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        train, momentum, eps)


def conv3_block(x, sd, p, train, stride=1):
    """basics.py:48-59 — (conv3x3 no-bias, BN, ReLU) twice."""
    x = _st(x)   # a block's input is a stored tensor (the network input is converted to the storage dtype at the module boundary)
    for a, b in (("0", "1"), ("3", "4")):
        x = _conv(x, sd[p + a + ".weight"], None, stride, 1)
        x = _st(torch.relu(batchnorm(x, sd, p + b + ".", train)))
    return x


def eca_kernel_size(channels, gamma=2, b=1):
    """basics.py:67-68."""
    t = int(abs((math.log2(channels) + b) / gamma))
    return t if t % 2 else t + 1


def eca(x, w):
    """basics.py:71-77 — channel attention: GAP -> conv1d over the channel axis -> sigmoid -> scale."""
    k = w.shape[-1]
    y = x.mean(dim=(2, 3))
    y = F.conv1d(y.unsqueeze(1), w, None, 1, k // 2).squeeze(1)
    return _st(x * torch.sigmoid(y)[:, :, None, None])


def eca_conv_block(x, sd, p, train, stride=1):
    """basics.py:80-135 — EfficientConvBlock."""
    x = eca(_st(x), sd[p + "layer1.eca1.conv.weight"])
    x = _conv(x, sd[p + "layer1.conv1.0.weight"], None, stride, 1)
    x = _st(torch.relu(batchnorm(x, sd, p + "layer1.conv1.1.", train)))
    x = eca(x, sd[p + "layer2.eca2.conv.weight"])
    x = _conv(x, sd[p + "layer2.conv2.0.weight"], None, stride, 1)
    return _st(torch.relu(batchnorm(x, sd, p + "layer2.conv2.1.", train)))


def mlp_layout(dims, act, l_act=False, bn=True, dropout=0.0):
    """basics.py:11-45 — the nn.Sequential index of every layer make_mlp creates."""
    ops, idx = [], 0
    n = len(dims) - 1
    for i in range(n):
        ops.append(("linear", idx, dims[i], dims[i + 1], not bn))
        idx += 1
        if i != n - 1:
            if bn:
                ops.append(("bn", idx, dims[i + 1]))
                idx += 1
            ops.append(("act", idx, act.lower()))
            idx += 1
            if dropout > 0.0:
                ops.append(("dropout", idx, dropout))
                idx += 1
    if l_act:
        ops.append(("act", idx, act.lower()))
    return ops


_ACTS = {"relu": torch.relu, "tanh": torch.tanh, "sigmoid": torch.sigmoid, "elu": F.elu}


def mlp(x, sd, p, cfg, train):
    """basics.py:11-45 forward of the Sequential built by make_mlp(**cfg)."""
    ops = mlp_layout(cfg["dims"], cfg["act"], cfg.get("l_act", False), cfg.get("bn", True), cfg.get("dropout", 0.0))
    for i, op in enumerate(ops):
        fused_act = i + 1 < len(ops) and ops[i + 1][0] == "act"  # the product applies bias + activation in the GEMM epilogue, then stores
        if op[0] == "linear":
            x = F.linear(x, _wq(sd[p + "%d.weight" % op[1]]), sd.get(p + "%d.bias" % op[1]))
            if not fused_act:
                x = _st(x)
        elif op[0] == "bn":
            x = batchnorm(x, sd, p + "%d." % op[1], train)
        elif op[0] == "act":
            x = _st(_ACTS[op[2]](x))
        elif op[0] == "dropout":
            x = F.dropout(x, op[2], train)
    return x


# ------------------------------------------------------------------ U-Net (model/blocks/unet.py:8-95)
def unet(x, sd, p, train, inter_repr=False):
    x1 = conv3_block(x, sd, p + "dwn_1.", train)
    x2 = conv3_block(F.max_pool2d(x1, 2, 2), sd, p + "dwn_2.", train)
    x3 = conv3_block(F.max_pool2d(x2, 2, 2), sd, p + "dwn_3.", train)
    x4 = conv3_block(F.max_pool2d(x3, 2, 2), sd, p + "dwn_4.", train)
    x5 = conv3_block(F.max_pool2d(x4, 2, 2), sd, p + "dwn_5.", train)
    y = x5
    for i, skip in ((1, x4), (2, x3), (3, x2), (4, x1)):
        up = _st(F.conv_transpose2d(y, _wq(sd[p + "up_%d.weight" % i]), sd[p + "up_%d.bias" % i], stride=2))
        # unet.py:72 output_size=skip.size(): pads bottom/right when the skip is odd-sized
        dh, dw = skip.shape[-2] - up.shape[-2], skip.shape[-1] - up.shape[-1]
        if dh or dw:
            up = F.conv_transpose2d(y, sd[p + "up_%d.weight" % i], sd[p + "up_%d.bias" % i], stride=2,
                                    output_padding=(dh, dw))
        y = conv3_block(torch.cat([skip, up], 1), sd, p + "up_forw_%d." % i, train)  # skip first (unet.py:73)
    out = _conv(y, sd[p + "out.weight"], sd[p + "out.bias"], 1, 0)
    if inter_repr:
        return x5.mean(dim=(2, 3)), out
    return out


def unet_eca(x, sd, p, train, inter_repr=False):
    """model/blocks/unet.py:98-185 — narrower U-Net (32..512) with an ECA gate on the pooled x_4 before dwn_5 (:160) and
    on every decoder concat before its conv3 (:165-178). NB inter_repr pools the POST-dwn_5 tensor (:182-184)."""
    x1 = conv3_block(x, sd, p + "dwn_1.", train)
    x2 = conv3_block(F.max_pool2d(x1, 2, 2), sd, p + "dwn_2.", train)
    x3 = conv3_block(F.max_pool2d(x2, 2, 2), sd, p + "dwn_3.", train)
    x4 = conv3_block(F.max_pool2d(x3, 2, 2), sd, p + "dwn_4.", train)
    x5 = conv3_block(eca(F.max_pool2d(x4, 2, 2), sd[p + "eca_0.conv.weight"]), sd, p + "dwn_5.", train)
    y = x5
    for i, skip in ((1, x4), (2, x3), (3, x2), (4, x1)):
        up = F.conv_transpose2d(y, sd[p + "up_%d.weight" % i], sd[p + "up_%d.bias" % i], stride=2)
        cat = torch.cat([skip, up], 1)
        y = conv3_block(eca(cat, sd[p + "eca_%d.conv.weight" % i]), sd, p + "up_forw_%d." % i, train)
    out = F.conv2d(y, sd[p + "out.weight"], sd[p + "out.bias"])
    if inter_repr:
        return x5.mean(dim=(2, 3)), out
    return out


# ------------------------------------------------------------------ ResNet18 + ECA stem (backbone.py:48-72)
RESNET18_LAYERS = ((64, 1), (128, 2), (256, 2), (512, 2))


RESNET_BLOCKS = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet34": ("basic", (3, 4, 6, 3)), "resnet50": ("bottleneck", (3, 4, 6, 3))}


def resnet_eca(x, sd, p, train, arch="resnet18"):
    """torchvision ResNet._forward_impl with conv1 := EfficientConvBlock and fc := Identity (512 features) or
    Linear(fc.in_features, 512) (backbone.py:48-72: resnet18/34 BasicBlock, resnet50 Bottleneck with the stride on the 3x3)."""
    kind, blocks = RESNET_BLOCKS[arch]
    x = eca_conv_block(x, sd, p + "conv1.", train)
    x = _st(torch.relu(batchnorm(x, sd, p + "bn1.", train)))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, ((_, stride), nb) in enumerate(zip(RESNET18_LAYERS, blocks), start=1):
        for bi in range(nb):
            q = p + "layer%d.%d." % (li, bi)
            s = stride if bi == 0 else 1
            idt = x
            if kind == "basic":
                y = _conv(x, sd[q + "conv1.weight"], None, s, 1)
                y = _st(torch.relu(batchnorm(y, sd, q + "bn1.", train)))
                y = _conv(y, sd[q + "conv2.weight"], None, 1, 1)
                y = batchnorm(y, sd, q + "bn2.", train)   # the residual add + ReLU ride the same pass: one stored tensor
            else:
                y = _st(torch.relu(batchnorm(_conv(x, sd[q + "conv1.weight"], None, 1, 0), sd, q + "bn1.", train)))
                y = _st(torch.relu(batchnorm(_conv(y, sd[q + "conv2.weight"], None, s, 1), sd, q + "bn2.", train)))
                y = batchnorm(_conv(y, sd[q + "conv3.weight"], None, 1, 0), sd, q + "bn3.", train)
            if q + "downsample.0.weight" in sd:
                idt = _conv(x, sd[q + "downsample.0.weight"], None, s, 0)
                idt = _st(batchnorm(idt, sd, q + "downsample.1.", train))
            x = _st(torch.relu(y + idt))
    x = _st(x.mean(dim=(2, 3)))
    if p + "fc.weight" in sd:
        x = _st(F.linear(x, _wq(sd[p + "fc.weight"]), sd[p + "fc.bias"]))
    return x


def resnet18_eca(x, sd, p, train):
    return resnet_eca(x, sd, p, train, "resnet18")


# ------------------------------------------------------------------ MobileNetV2 + ECA stem (backbone.py:75-104)
# torchvision MobileNetV2 (pinned 0.9.1, unchanged arithmetic in 0.26): features[0] = (Conv3x3 s2 -> replaced by the reference with
# a STRIDE-1 EfficientConvBlock(n_frames*3 -> 32), BN, ReLU6); 17 InvertedResidual blocks (t, c, n, s below: 1x1 expand + BN +
# ReLU6 unless t == 1, depthwise 3x3 stride s + BN + ReLU6, 1x1 project + BN, identity add when s == 1 and cin == cout);
# features[18] = 1x1 conv 320 -> 1280 + BN + ReLU6; global average pool; classifier := Linear(1280, 512) (backbone.py:98-99).
# Oracle groundwork for SURVEY §8 row a8 — the product does not build this family yet (get_backbone raises for it).
MOBILENET_V2_CFG = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))


def _mbv2_blocks(cin=32):
    out, idx = [], 1
    for t, c, n, s in MOBILENET_V2_CFG:
        for i in range(n):
            out.append((idx, cin, c, s if i == 0 else 1, t))
            cin, idx = c, idx + 1
    return out


def mobilenet_v2_eca(x, sd, p, train):
    x = eca_conv_block(x, sd, p + "features.0.0.", train)
    x = F.relu6(batchnorm(x, sd, p + "features.0.1.", train))
    for idx, cin, cout, stride, t in _mbv2_blocks():
        q = p + "features.%d.conv." % idx
        y, j = x, 0
        if t != 1:
            y = F.relu6(batchnorm(F.conv2d(y, sd[q + "0.0.weight"]), sd, q + "0.1.", train))
            j = 1
        hid = y.shape[1]
        y = F.relu6(batchnorm(F.conv2d(y, sd[q + "%d.0.weight" % j], None, stride, 1, 1, hid), sd, q + "%d.1." % j, train))
        y = batchnorm(F.conv2d(y, sd[q + "%d.weight" % (j + 1)]), sd, q + "%d." % (j + 2), train)
        x = x + y if (stride == 1 and cin == cout) else y
    x = F.relu6(batchnorm(F.conv2d(x, sd[p + "features.18.0.weight"]), sd, p + "features.18.1.", train))
    x = x.mean(dim=(2, 3))
    return F.linear(x, sd[p + "classifier.weight"], sd[p + "classifier.bias"])


# torchvision MobileNetV3-Small (the arch `_get_mobilenet` falls back to, backbone.py:86-90): (cin, kernel, expanded, cout, SE, act, stride)
MOBILENET_V3_SMALL_CFG = ((16, 3, 16, 16, True, "RE", 2), (16, 3, 72, 24, False, "RE", 2), (24, 3, 88, 24, False, "RE", 1),
                          (24, 5, 96, 40, True, "HS", 2), (40, 5, 240, 40, True, "HS", 1), (40, 5, 240, 40, True, "HS", 1),
                          (40, 5, 120, 48, True, "HS", 1), (48, 5, 144, 48, True, "HS", 1), (48, 5, 288, 96, True, "HS", 2),
                          (96, 5, 576, 96, True, "HS", 1), (96, 5, 576, 96, True, "HS", 1))
MOBILENET_V3_LARGE_CFG = ((16, 3, 16, 16, False, "RE", 1), (16, 3, 64, 24, False, "RE", 2), (24, 3, 72, 24, False, "RE", 1),
                          (24, 5, 72, 40, True, "RE", 2), (40, 5, 120, 40, True, "RE", 1), (40, 5, 120, 40, True, "RE", 1),
                          (40, 3, 240, 80, False, "HS", 2), (80, 3, 200, 80, False, "HS", 1), (80, 3, 184, 80, False, "HS", 1),
                          (80, 3, 184, 80, False, "HS", 1), (80, 3, 480, 112, True, "HS", 1), (112, 3, 672, 112, True, "HS", 1),
                          (112, 5, 672, 160, True, "HS", 2), (160, 5, 960, 160, True, "HS", 1), (160, 5, 960, 160, True, "HS", 1))
MOBILENET_V3 = {"mobilenet_v3_small": (MOBILENET_V3_SMALL_CFG, 1024), "mobilenet_v3_large": (MOBILENET_V3_LARGE_CFG, 1280)}
MBV3_BN = dict(eps=0.001, momentum=0.01)   # norm_layer = partial(BatchNorm2d, eps=0.001, momentum=0.01)


def _make_divisible(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    return new_v + divisor if new_v < 0.9 * v else new_v


def mobilenet_v3_small_eca(x, sd, p, train):
    return mobilenet_v3_eca(x, sd, p, train, "mobilenet_v3_small")


def mobilenet_v3_large_eca(x, sd, p, train):
    return mobilenet_v3_eca(x, sd, p, train, "mobilenet_v3_large")


def mobilenet_v3_eca(x, sd, p, train, arch="mobilenet_v3_small"):
    """Stride-1 ECA stem (16 channels) + BN + Hardswish; 11 inverted-residual blocks (1x1 expand unless expanded == cin, depthwise
    k x k, optional squeeze-excite with ReLU / Hardsigmoid, 1x1 project, identity add when stride 1 and cin == cout); 1x1 conv to 576
    + BN + Hardswish; global average pool; Linear(576, 1024) + Hardswish (+ Dropout, identity here) + Linear(1024, 512)."""
    x = eca_conv_block(x, sd, p + "features.0.0.", train)
    x = F.hardswish(batchnorm(x, sd, p + "features.0.1.", train, **MBV3_BN))
    cfg, _ = MOBILENET_V3[arch]
    last = len(cfg) + 1
    for i, (cin, k, exp, cout, se, act, stride) in enumerate(cfg, start=1):
        fn = F.hardswish if act == "HS" else F.relu
        q = p + "features.%d.block." % i
        y, j = x, 0
        if exp != cin:
            y = fn(batchnorm(F.conv2d(y, sd[q + "0.0.weight"]), sd, q + "0.1.", train, **MBV3_BN))
            j = 1
        y = fn(batchnorm(F.conv2d(y, sd[q + "%d.0.weight" % j], None, stride, (k - 1) // 2, 1, exp), sd, q + "%d.1." % j, train, **MBV3_BN))
        j += 1
        if se:
            z = y.mean(dim=(2, 3), keepdim=True)
            z = F.relu(F.conv2d(z, sd[q + "%d.fc1.weight" % j], sd[q + "%d.fc1.bias" % j]))
            z = F.hardsigmoid(F.conv2d(z, sd[q + "%d.fc2.weight" % j], sd[q + "%d.fc2.bias" % j]))
            y = y * z
            j += 1
        y = batchnorm(F.conv2d(y, sd[q + "%d.0.weight" % j]), sd, q + "%d.1." % j, train, **MBV3_BN)
        x = x + y if (stride == 1 and cin == cout) else y
    x = F.hardswish(batchnorm(F.conv2d(x, sd[p + "features.%d.0.weight" % last]), sd, p + "features.%d.1." % last, train, **MBV3_BN))
    x = x.mean(dim=(2, 3))
    x = F.hardswish(F.linear(x, sd[p + "classifier.0.weight"], sd[p + "classifier.0.bias"]))
    return F.linear(x, sd[p + "classifier.3.weight"], sd[p + "classifier.3.bias"])


def mobilenet_v3_small_spec(spec, p, cin, gamma=2, b=1):
    mobilenet_v3_spec(spec, p, cin, gamma, b, "mobilenet_v3_small")


def mobilenet_v3_large_spec(spec, p, cin, gamma=2, b=1):
    mobilenet_v3_spec(spec, p, cin, gamma, b, "mobilenet_v3_large")


def mobilenet_v3_spec(spec, p, cin, gamma=2, b=1, arch="mobilenet_v3_small"):
    cfg, hidden = MOBILENET_V3[arch]
    last, c_last = len(cfg) + 1, cfg[-1][3]
    eca_block_spec(spec, p + "features.0.0.", cin, 16, gamma, b)
    _bn(spec, p + "features.0.1.", 16)
    for i, (ci, k, exp, co, se, act, stride) in enumerate(cfg, start=1):
        q = p + "features.%d.block." % i
        j = 0
        if exp != ci:
            spec[q + "0.0.weight"] = (exp, ci, 1, 1)
            _bn(spec, q + "0.1.", exp)
            j = 1
        spec[q + "%d.0.weight" % j] = (exp, 1, k, k)
        _bn(spec, q + "%d.1." % j, exp)
        j += 1
        if se:
            sq = _make_divisible(exp // 4, 8)
            spec[q + "%d.fc1.weight" % j], spec[q + "%d.fc1.bias" % j] = (sq, exp, 1, 1), (sq,)
            spec[q + "%d.fc2.weight" % j], spec[q + "%d.fc2.bias" % j] = (exp, sq, 1, 1), (exp,)
            j += 1
        spec[q + "%d.0.weight" % j] = (co, exp, 1, 1)
        _bn(spec, q + "%d.1." % j, co)
    spec[p + "features.%d.0.weight" % last] = (6 * c_last, c_last, 1, 1)
    _bn(spec, p + "features.%d.1." % last, 6 * c_last)
    spec[p + "classifier.0.weight"], spec[p + "classifier.0.bias"] = (hidden, 6 * c_last), (hidden,)
    spec[p + "classifier.3.weight"], spec[p + "classifier.3.bias"] = (512, hidden), (512,)


def mobilenet_v2_spec(spec, p, cin, gamma=2, b=1):
    eca_block_spec(spec, p + "features.0.0.", cin, 32, gamma, b)
    _bn(spec, p + "features.0.1.", 32)
    for idx, ci, co, stride, t in _mbv2_blocks():
        q = p + "features.%d.conv." % idx
        hid, j = ci * t, 0
        if t != 1:
            spec[q + "0.0.weight"] = (hid, ci, 1, 1)
            _bn(spec, q + "0.1.", hid)
            j = 1
        spec[q + "%d.0.weight" % j] = (hid, 1, 3, 3)
        _bn(spec, q + "%d.1." % j, hid)
        spec[q + "%d.weight" % (j + 1)] = (co, hid, 1, 1)
        _bn(spec, q + "%d." % (j + 2), co)
    spec[p + "features.18.0.weight"] = (1280, 320, 1, 1)
    _bn(spec, p + "features.18.1.", 1280)
    spec[p + "classifier.weight"] = (512, 1280)
    spec[p + "classifier.bias"] = (512,)


# ------------------------------------------------------------------ PU-Net (model/punet.py:75-120)
def punet(imgs, sd, p, train, past_frames=4, future_frames=6, inter_repr=False, unet_inter_repr=False):
    assert imgs.shape[-4] == past_frames
    masks = [unet(imgs[:, i], sd, p + "unet.", train, unet_inter_repr) for i in range(past_frames)]
    if future_frames == 0:
        return masks[-1][0] if unet_inter_repr else masks[-1]
    outs, inter = [], None
    for _ in range(future_frames):
        m = torch.cat(masks[-past_frames:], dim=-3)
        m = eca_conv_block(m, sd, p + "entry_block.", train)
        if inter_repr:
            inter, m = unet(m, sd, p + "pred_unet.", train, True)
        else:
            m = unet(m, sd, p + "pred_unet.", train, False)
        masks.append(m)
        outs.append(m)
    return inter if inter_repr else torch.stack(outs, dim=1)


# ------------------------------------------------------------------ experts / mixtures (model/moe.py)
def expert(images, speed, command, sd, p, cfg, train, alt=False):
    """moe.py:74-101 (BaseExpert) / :113-128 (BaseExpertAlt) -> alpha, mean, std, pred_speed."""
    s = mlp(speed, sd, p + "speed_encoder.", cfg["speed_encoder"], train)
    c = mlp(command, sd, p + "command_encoder.", cfg["command_encoder"], train)
    img = resnet18_eca(images.reshape(images.shape[0], -1, images.shape[-2], images.shape[-1]), sd, p + "backbone.", train)
    feats = torch.cat([img, s, c], dim=-1)
    pred_speed = mlp(feats, sd, p + "speed_pred.", cfg["speed_prediction"], train)
    af = mlp(feats, sd, p + "action_features.", cfg["action_head"], train)
    mean, std = _st(F.linear(af, _wq(sd[p + "action_pred.weight"]), sd[p + "action_pred.bias"])).split(2, dim=-1)
    std = F.elu(std) + 1
    if alt:
        a = F.linear(feats, _wq(sd[p + "alpha.0.weight"]), sd[p + "alpha.0.bias"])
        alpha = _st(F.linear(_st(torch.relu(a)), _wq(sd[p + "alpha.2.weight"]), sd[p + "alpha.2.bias"]))
    else:
        alpha = torch.relu(_st(F.linear(af, _wq(sd[p + "alpha.weight"]), sd[p + "alpha.bias"])))  # the gating kernel applies the ReLU
    return alpha, mean, std, pred_speed


def moe(images, speed, command, sd, p, cfg, train):
    """moe.py:140-158 -> (probs (B,K), mean (B,K,2), std (B,K,2), speeds (B,K,1))."""
    # moe.py:136: every type other than 'moe' (moe_alt, and PMoE's inner mixture) builds BaseExpertAlt
    outs = [expert(images, speed, command, sd, p + "moe.%d." % k, cfg, train, cfg["type"] != "moe")
            for k in range(cfg["n_experts"])]
    probs = F.softmax(torch.cat([o[0] for o in outs], dim=1), dim=1)
    return probs, torch.stack([o[1] for o in outs], 1), torch.stack([o[2] for o in outs], 1), torch.stack([o[3] for o in outs], 1)


def moe_shared(images, speed, command, sd, p, cfg, train):
    """moe.py:205-239 -> (probs (B,K), mean, std, pred_speed (B,1))."""
    K = cfg["n_experts"]
    s = mlp(speed, sd, p + "speed_encoder.", cfg["speed_encoder"], train)
    c = mlp(command, sd, p + "command_encoder.", cfg["command_encoder"], train)
    img = resnet18_eca(images.reshape(images.shape[0], -1, images.shape[-2], images.shape[-1]), sd, p + "backbone.", train)
    feats = torch.cat([img, s, c], dim=-1)
    pred_speed = mlp(feats, sd, p + "speed_pred.", cfg["speed_prediction"], train)
    af = mlp(feats, sd, p + "action_features.", cfg["action_head"], train)
    mean, std = F.linear(af, sd[p + "action_pred.weight"], sd[p + "action_pred.bias"]).view(af.shape[0], K, -1).split(2, dim=-1)
    std = F.elu(std) + 1
    probs = F.softmax(F.linear(af, sd[p + "alpha.weight"], sd[p + "alpha.bias"]), dim=1)
    return probs, mean, std, pred_speed


def mixture_log_prob(probs, mean, std, x):
    """torch.distributions.MixtureSameFamily(Categorical(probs), Independent(Normal(mean,std),1)).log_prob
    as called at trainer/loss.py:123 (torch 2.11 semantics: Categorical renormalises probs and uses
    logits = log(clamp(p, eps, 1-eps)); log_softmax of those logits is taken again)."""
    eps = torch.finfo(probs.dtype).eps
    pn = probs / probs.sum(-1, keepdim=True)
    logits = torch.log(pn.clamp(eps, 1 - eps))
    log_mix = torch.log_softmax(logits, dim=-1)
    xe = x.unsqueeze(-2)
    comp = (-((xe - mean) ** 2) / (2 * std ** 2) - std.log() - 0.5 * math.log(2 * math.pi)).sum(-1)
    return torch.logsumexp(comp + log_mix, dim=-1)


def mixture_sample(probs, mean, std, generator=None):
    """MixtureSameFamily.sample() (moe.py:177,352): Categorical index via multinomial, then a full
    Normal draw of every component, gathered. RNG consumption order matches torch.distributions."""
    idx = torch.multinomial(probs / probs.sum(-1, keepdim=True), 1, True, generator=generator)  # (B,1)
    comp = torch.normal(mean, std, generator=generator)  # (B,K,2)
    return comp.gather(1, idx.unsqueeze(-1).expand(-1, 1, comp.shape[-1])).squeeze(1), idx.squeeze(1)


def punet_expert(images, speed, command, sd, p, cfg, train):
    """moe.py:303-317 -> (tanh(action) (B,2), speed (B,1))."""
    inter = cfg["type"] == "punet_inter"
    pc = cfg["punet"]
    s = mlp(speed, sd, p + "speed_encoder.", cfg["speed_encoder"], train)
    c = mlp(command, sd, p + "command_encoder.", cfg["command_encoder"], train)
    out = punet(images, sd, p + "punet.", train, pc["past_frames"], pc["future_frames"], inter, pc.get("unet_inter_repr", False))
    if inter:
        img = out
    else:
        img = resnet18_eca(out.reshape(out.shape[0], -1, out.shape[-2], out.shape[-1]), sd, p + "backbone.", train)
    feats = torch.cat([img, s, c], dim=-1)
    a = mlp(feats, sd, p + "action_pred.0.", cfg["action_head"], train)
    a = F.linear(a, sd[p + "action_pred.1.weight"], sd[p + "action_pred.1.bias"])
    return torch.tanh(a), mlp(feats, sd, p + "speed_pred.", cfg["speed_prediction"], train)


def pmoe_combine(moe_actions, punet_actions, sd, p):
    """moe.py:353-356."""
    lat = F.linear(torch.cat([moe_actions[:, 0:1], punet_actions[:, 0:1]], -1), sd[p + "lat_weights.weight"], sd[p + "lat_weights.bias"])
    lon = F.linear(torch.cat([moe_actions[:, 1:], punet_actions[:, 1:]], -1), sd[p + "long_weights.weight"], sd[p + "long_weights.bias"])
    return torch.tanh(torch.cat([lat, lon], -1))


# ------------------------------------------------------------------ losses (trainer/loss.py)
def class_dice_weights(pred, target, eps=1e-6):
    """loss.py:6-17 — 1 - dice per class from argmax predictions (no gradient)."""
    nc = pred.shape[1]
    pc = pred.argmax(1)
    w = torch.ones(nc, dtype=pred.dtype)
    for c in range(nc):
        pm, tm = pc == c, target == c
        inter = (pm & tm).sum().float() + eps
        union = pm.sum() + tm.sum() + eps
        w[c] = 1 - 2 * inter / union
    return w


def tversky(pred, target, alpha=0.5, beta=0.5):
    """loss.py:34-44."""
    oh = F.one_hot(target, pred.shape[1]).movedim(-1, 1).to(pred.dtype)
    pr = F.softmax(pred, 1)
    # loss.py:40: range(2, target.ndimension()) with a (B,H,W) target is (0, 2) only — the sums run over
    # batch and HEIGHT, leaving a (C, W) ratio map that is then averaged. Reproduced, not "fixed".
    dims = (0,) + tuple(range(2, target.dim()))
    tp = (pr * oh).sum(dims)
    fp = (pr * (1 - oh)).sum(dims)
    fn = ((1 - pr) * oh).sum(dims)
    return 1 - (tp / (tp + alpha * fp + beta * fn)).mean()


def ce_tversky(pred, target, wc=0.5, wt=0.5):
    """loss.py:47-55."""
    ce = F.cross_entropy(pred, target, weight=class_dice_weights(pred, target))
    return wc * ce + wt * tversky(pred, target)


def dice_score(pred, target, eps=1e-6):
    """loss.py:20-31 — per-class dice of the argmax prediction (validation metric; = 1 - class_dice weights)."""
    return 1 - class_dice_weights(pred, target, eps)


def l1_gdl(inputs, targets):
    """loss.py:58-83 — L1 + gradient-difference loss on the LAST frame: raw logits against the one-hot target (the softmax
    at :66 is computed and never used). Differences run against a zero row/column appended at the bottom/right (:67-68);
    the gdl term is summed over (H, W) and averaged over (B, C) (:79), the L1 term is a plain mean (:81)."""
    x = inputs[:, -1]
    oh = F.one_hot(targets[:, -1], x.shape[1]).movedim(-1, 1).to(x.dtype)

    def dv(t):
        tp = F.pad(t, (0, 0, 0, 1))
        return (tp[..., 1:, :] - tp[..., :-1, :]).abs()

    def dh(t):
        tp = F.pad(t, (0, 1, 0, 0))
        return (tp[..., :-1] - tp[..., 1:]).abs()

    gdl = ((dv(oh) - dv(x)).abs() + (dh(oh) - dh(x)).abs()).sum(dim=(-2, -1)).mean()
    return (x - oh).abs().mean() + gdl


def autoregressive_onehot(inputs, targets, loss_type):
    """loss.py:86-118 with loss_type 'l1' / 'l2': per-frame nn.L1Loss / nn.MSELoss against the one-hot targets."""
    oh = F.one_hot(targets, inputs.shape[2]).movedim(-1, 2).to(inputs.dtype)
    fn = F.l1_loss if loss_type == "l1" else F.mse_loss
    return sum(fn(inputs[:, t], oh[:, t]) for t in range(inputs.shape[1]))


def autoregressive_ce_tversky(inputs, targets):
    """loss.py:86-118 with loss_type='tversky'."""
    return sum(ce_tversky(inputs[:, t], targets[:, t]) for t in range(inputs.shape[1]))


def moe_loss(probs, mean, std, speed_pred, actions_gt, speed_gt, coefs):
    """loss.py:121-132 (speed_gt is (B,1); the reference unsqueezes it in place when speed_pred is 3-D)."""
    nll = -mixture_log_prob(probs, mean, std, actions_gt).mean(0)
    if speed_pred.dim() > 2:
        sp = F.mse_loss(speed_pred, speed_gt.unsqueeze(1).expand_as(speed_pred)) / speed_pred.shape[1]
    else:
        sp = F.mse_loss(speed_pred, speed_gt)
    return coefs[0] * nll + coefs[1] * sp


def punet_loss(actions, speed_pred, actions_gt, speed_gt, coefs):
    """loss.py:135-142."""
    return coefs[0] * F.l1_loss(actions, actions_gt) + coefs[1] * F.mse_loss(speed_pred, speed_gt)


def pmoe_loss(actions, actions_gt):
    """loss.py:145-151."""
    return F.l1_loss(actions, actions_gt)


# ------------------------------------------------------------------ state-dict specs
def _bn(spec, p, c):
    spec[p + "weight"] = (c,)
    spec[p + "bias"] = (c,)
    spec[p + "running_mean"] = (c,)
    spec[p + "running_var"] = (c,)
    spec[p + "num_batches_tracked"] = ()


def conv3_spec(spec, p, cin, cout):
    spec[p + "0.weight"] = (cout, cin, 3, 3)
    _bn(spec, p + "1.", cout)
    spec[p + "3.weight"] = (cout, cout, 3, 3)
    _bn(spec, p + "4.", cout)


def eca_block_spec(spec, p, cin, cout, gamma=2, b=1):
    spec[p + "layer1.eca1.conv.weight"] = (1, 1, eca_kernel_size(cin, gamma, b))
    spec[p + "layer1.conv1.0.weight"] = (64, cin, 3, 3)
    _bn(spec, p + "layer1.conv1.1.", 64)
    spec[p + "layer2.eca2.conv.weight"] = (1, 1, eca_kernel_size(64, gamma, b))
    spec[p + "layer2.conv2.0.weight"] = (cout, 64, 3, 3)
    _bn(spec, p + "layer2.conv2.1.", cout)


def unet_spec(spec, p, cin=3, cout=23):
    for i, (a, b) in enumerate(((cin, 64), (64, 128), (128, 256), (256, 512), (512, 512)), start=1):
        conv3_spec(spec, p + "dwn_%d." % i, a, b)
    for i, (a, b) in enumerate(((512, 512), (512, 256), (256, 128), (128, 64)), start=1):
        spec[p + "up_%d.weight" % i] = (a, b, 2, 2)
        spec[p + "up_%d.bias" % i] = (b,)
        conv3_spec(spec, p + "up_forw_%d." % i, 2 * b, b)
    spec[p + "out.weight"] = (cout, 64, 1, 1)
    spec[p + "out.bias"] = (cout,)


def unet_eca_spec(spec, p, cin=3, cout=23, gamma=2, b=1):
    """unet.py:111-138 (module registration order)."""
    for i, (a, c) in enumerate(((cin, 32), (32, 64), (64, 128), (128, 256), (256, 512)), start=1):
        conv3_spec(spec, p + "dwn_%d." % i, a, c)
    spec[p + "eca_0.conv.weight"] = (1, 1, eca_kernel_size(512, gamma, b))
    for i, (a, c) in enumerate(((512, 256), (256, 128), (128, 64), (64, 32)), start=1):
        spec[p + "up_%d.weight" % i] = (a, c, 2, 2)
        spec[p + "up_%d.bias" % i] = (c,)
        spec[p + "eca_%d.conv.weight" % i] = (1, 1, eca_kernel_size(2 * c, gamma, b))
        conv3_spec(spec, p + "up_forw_%d." % i, 2 * c, c)
    spec[p + "out.weight"] = (cout, 32, 1, 1)
    spec[p + "out.bias"] = (cout,)


def resnet_spec(spec, p, cin, gamma=2, b=1, arch="resnet18"):
    kind, blocks = RESNET_BLOCKS[arch]
    exp = 1 if kind == "basic" else 4
    eca_block_spec(spec, p + "conv1.", cin, 64, gamma, b)
    _bn(spec, p + "bn1.", 64)
    prev = 64
    for li, ((c, stride), nb) in enumerate(zip(RESNET18_LAYERS, blocks), start=1):
        for bi in range(nb):
            q = p + "layer%d.%d." % (li, bi)
            if kind == "basic":
                spec[q + "conv1.weight"] = (c, prev, 3, 3)
                _bn(spec, q + "bn1.", c)
                spec[q + "conv2.weight"] = (c, c, 3, 3)
                _bn(spec, q + "bn2.", c)
            else:
                spec[q + "conv1.weight"] = (c, prev, 1, 1)
                _bn(spec, q + "bn1.", c)
                spec[q + "conv2.weight"] = (c, c, 3, 3)
                _bn(spec, q + "bn2.", c)
                spec[q + "conv3.weight"] = (c * exp, c, 1, 1)
                _bn(spec, q + "bn3.", c * exp)
            if bi == 0 and (stride != 1 or prev != c * exp):
                spec[q + "downsample.0.weight"] = (c * exp, prev, 1, 1)
                _bn(spec, q + "downsample.1.", c * exp)
            prev = c * exp
    if prev != 512:
        spec[p + "fc.weight"] = (512, prev)
        spec[p + "fc.bias"] = (512,)


def resnet18_spec(spec, p, cin, gamma=2, b=1):
    resnet_spec(spec, p, cin, gamma, b, "resnet18")


def mlp_spec(spec, p, cfg):
    for op in mlp_layout(cfg["dims"], cfg["act"], cfg.get("l_act", False), cfg.get("bn", True), cfg.get("dropout", 0.0)):
        if op[0] == "linear":
            spec[p + "%d.weight" % op[1]] = (op[3], op[2])
            if op[4]:
                spec[p + "%d.bias" % op[1]] = (op[3],)
        elif op[0] == "bn":
            _bn(spec, p + "%d." % op[1], op[2])


def expert_spec(spec, p, cfg, alt=False):
    mlp_spec(spec, p + "speed_encoder.", cfg["speed_encoder"])
    mlp_spec(spec, p + "command_encoder.", cfg["command_encoder"])
    rgb = cfg["backbone"]["rgb"]
    resnet18_spec(spec, p + "backbone.", cfg["backbone"]["n_frames"] * 3, rgb.get("gamma", 2), rgb.get("b", 1))
    mlp_spec(spec, p + "speed_pred.", cfg["speed_prediction"])
    mlp_spec(spec, p + "action_features.", cfg["action_head"])
    d = cfg["action_head"]["dims"][-1]
    if alt:
        spec[p + "alpha.0.weight"], spec[p + "alpha.0.bias"] = (512, 1536), (512,)
        spec[p + "alpha.2.weight"], spec[p + "alpha.2.bias"] = (1, 512), (1,)
    else:
        spec[p + "alpha.weight"], spec[p + "alpha.bias"] = (1, d), (1,)
    spec[p + "action_pred.weight"], spec[p + "action_pred.bias"] = (4, d), (4,)


def moe_spec(spec, p, cfg):
    for k in range(cfg["n_experts"]):
        expert_spec(spec, p + "moe.%d." % k, cfg, cfg["type"] != "moe")


def moe_shared_spec(spec, p, cfg):
    K = cfg["n_experts"]
    mlp_spec(spec, p + "speed_encoder.", cfg["speed_encoder"])
    mlp_spec(spec, p + "command_encoder.", cfg["command_encoder"])
    rgb = cfg["backbone"]["rgb"]
    resnet18_spec(spec, p + "backbone.", cfg["backbone"]["n_frames"] * 3, rgb.get("gamma", 2), rgb.get("b", 1))
    mlp_spec(spec, p + "speed_pred.", cfg["speed_prediction"])
    mlp_spec(spec, p + "action_features.", cfg["action_head"])
    d = cfg["action_head"]["dims"][-1]
    spec[p + "alpha.weight"], spec[p + "alpha.bias"] = (K, d), (K,)
    spec[p + "action_pred.weight"], spec[p + "action_pred.bias"] = (4 * K, d), (4 * K,)


def punet_spec(spec, p, pc):
    unet_spec(spec, p + "unet.", pc["in_features"], pc["num_classes"])
    eca_block_spec(spec, p + "entry_block.", pc["past_frames"] * pc["num_classes"], pc["in_features"], pc.get("gamma", 2), pc.get("b", 1))
    unet_spec(spec, p + "pred_unet.", pc["in_features"], pc["num_classes"])


def punet_expert_spec(spec, p, cfg):
    pc = cfg["punet"]
    mlp_spec(spec, p + "speed_encoder.", cfg["speed_encoder"])
    mlp_spec(spec, p + "command_encoder.", cfg["command_encoder"])
    punet_spec(spec, p + "punet.", pc)
    if cfg["type"] != "punet_inter":
        rgb = cfg["backbone"]["rgb"]
        resnet18_spec(spec, p + "backbone.", pc["future_frames"] * pc["num_classes"], rgb.get("gamma", 2), rgb.get("b", 1))
    mlp_spec(spec, p + "speed_pred.", cfg["speed_prediction"])
    mlp_spec(spec, p + "action_pred.0.", cfg["action_head"])
    spec[p + "action_pred.1.weight"], spec[p + "action_pred.1.bias"] = (2, cfg["action_head"]["dims"][-1]), (2,)


def pmoe_spec(spec, p, cfg):
    moe_spec(spec, p + "moe.", cfg)
    punet_expert_spec(spec, p + "punet.", cfg)
    for n in ("lat_weights", "long_weights"):
        spec[p + n + ".weight"], spec[p + n + ".bias"] = (1, 2), (1,)


def make_spec(fn, *args):
    spec = OrderedDict()
    fn(spec, "", *args)
    return spec


def seeded_state_dict(spec, seed, w_gain=1.0):
    """Deterministic, well-conditioned weights for a spec (He-scaled convs/linears, non-trivial BN
    affine + running stats) — used by both the golden generator and the tests, so the big weight
    tensors never need to be stored."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in spec.items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.zeros((), dtype=torch.int64)
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "weight" and len(shape) == 1:
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.05
        elif leaf == "weight" and len(shape) == 3:  # ECA conv1d
            t = torch.randn(shape, generator=g) * 0.5
        else:
            if name.startswith("up_") or ".up_" in name:
                fan_in = shape[0]  # ConvTranspose2d k2s2: each output pixel sees Cin inputs
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
            t = torch.randn(shape, generator=g) * (w_gain * (2.0 / max(fan_in, 1)) ** 0.5)
        sd[name] = t
    return sd


DEFAULT_MODEL_CFG = {
    # conf/stage_2.yaml:76-133 with pretrained=False and dropout disabled for parity runs
    "type": "moe", "n_experts": 3, "loss_coefs": [0.7, 0.3], "exclude_freeze": [], "verbose": False,
    "action_head": {"dims": [1536, 512, 512], "act": "elu", "l_act": True, "bn": False, "dropout": 0.0},
    "speed_encoder": {"dims": [1, 512, 512], "act": "relu", "l_act": False, "bn": False, "dropout": 0.0},
    "command_encoder": {"dims": [6, 512, 512], "act": "relu", "l_act": False, "bn": False, "dropout": 0.0},
    "speed_prediction": {"dims": [1536, 512, 512, 1], "act": "relu", "l_act": False, "bn": False, "dropout": 0.0},
    "backbone": {"type": "rgb", "n_frames": 4, "rgb": {"arch": "resnet18", "pretrained": False, "gamma": 2, "b": 1}},
    "punet": {"past_frames": 4, "future_frames": 6, "in_features": 3, "num_classes": 23, "gamma": 2, "b": 1,
              "unet_inter_repr": False, "model_name": "unet", "model_path": ""},
    "pmoe": {"moe_dir": "", "punet_dir": ""}, "punet_path": "", "device": "cpu",
}
