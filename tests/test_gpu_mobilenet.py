"""the MobileNet family of the backbone factory (reference: model/blocks/backbone.py:75-104) on the GPU
kernels. Kernel level: depthwise convolution forward / data gradient / weight gradient and the ReLU6 / Hardswish / Hardsigmoid
forward + backward against ATen on the same operands. Model level: MobileNetV2 / V3-Small / V3-Large with the ECA stem against
the live-reference goldens (tests/golden/backbone_mobilenet_*.pt) — fp32 parity mode: eval and train features 1e-4, every
gradient as close to an fp64 evaluation as the reference's own fp32 run is; bf16: features, finite gradients."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_err
from oracle import functional as O

pytestmark = pytest.mark.gpu
dev = "cuda"


def _nhwc(t, dtype):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype)


@pytest.mark.parametrize("k,stride,h,w,c", [(3, 1, 14, 18, 32), (3, 2, 15, 17, 16), (5, 1, 12, 12, 72), (5, 2, 16, 20, 96), (3, 2, 8, 8, 960)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_depthwise_conv_kernels_vs_aten(k, stride, h, w, c, dtype):
    from pmoe_b200 import _lib
    from pmoe_b200._lib import check, lib, stream_ptr, view4
    g = torch.Generator().manual_seed(k * 100 + stride * 10 + c)
    n, pad = 3, (k - 1) // 2
    x = torch.randn(n, c, h, w, generator=g).to(dev)
    wt = torch.randn(c, 1, k, k, generator=g).to(dev) * 0.3
    xq = x.to(dtype).float()                     # operands as stored
    y_ref = F.conv2d(xq, wt, None, stride, pad, 1, c)
    oh, ow = y_ref.shape[2], y_ref.shape[3]
    dy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(7)).to(dev).to(dtype).float()
    xg = xq.clone().requires_grad_(True)
    wg = wt.clone().requires_grad_(True)
    F.conv2d(xg, wg, None, stride, pad, 1, c).backward(dy)
    code = _lib.BF16 if dtype == torch.bfloat16 else _lib.F32
    wp = wt[:, 0].permute(1, 2, 0).reshape(k * k, c).contiguous()
    xn, dyn = _nhwc(xq, dtype), _nhwc(dy, dtype)
    yn = torch.empty(n, oh, ow, c, dtype=dtype, device=dev)
    vx, vy, vdy = view4(xn), view4(yn), view4(dyn)
    check(lib().pmoe_dwconv_fwd(C.byref(vx), wp.data_ptr(), c, C.byref(vy), code, k, stride, pad, stream_ptr()), "fwd")
    tol = 1e-5 if dtype == torch.float32 else 6e-3
    assert rel_err(yn.float().permute(0, 3, 1, 2), y_ref) < tol
    dxn = torch.empty_like(xn)
    vdx = view4(dxn)
    check(lib().pmoe_dwconv_dgrad(C.byref(vdy), wp.data_ptr(), c, C.byref(vdx), code, k, stride, pad, 0, stream_ptr()), "dgrad")
    assert rel_err(dxn.float().permute(0, 3, 1, 2), xg.grad) < tol
    check(lib().pmoe_dwconv_dgrad(C.byref(vdy), wp.data_ptr(), c, C.byref(vdx), code, k, stride, pad, 1, stream_ptr()), "dgrad acc")
    assert rel_err(dxn.float().permute(0, 3, 1, 2), 2 * xg.grad) < 2 * tol
    dwp = torch.zeros(k * k, c, dtype=torch.float32, device=dev)
    check(lib().pmoe_dwconv_wgrad(C.byref(vx), C.byref(vdy), dwp.data_ptr(), c, code, k, stride, pad, stream_ptr()), "wgrad")
    assert rel_err(dwp.view(k, k, c).permute(2, 0, 1), wg.grad[:, 0]) < 1e-5   # fp32 accumulation of exactly the stored operands


@pytest.mark.parametrize("act,fn", [("relu6", F.relu6), ("hswish", F.hardswish), ("hsigmoid", F.hardsigmoid)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mobilenet_activations_forward_and_backward_vs_aten(act, fn, dtype):
    """act_op: y = act(x) through the affine kernel, dx = dy * act'(x) through the BatchNorm-backward kernel's pre-activation path."""
    from pmoe_b200 import config, nhwc, train
    x = (torch.randn(2, 9, 11, 16, generator=torch.Generator().manual_seed(3)) * 4).to(dev).to(dtype)
    x[0, 0, 0, :4] = torch.tensor([-3.0, 3.0, 0.0, 6.0], device=dev).to(dtype)    # the kinks themselves
    dy = torch.randn(x.shape, generator=torch.Generator().manual_seed(4)).to(dev).to(dtype)
    xr = x.float().clone().requires_grad_(True)
    yr = fn(xr)
    yr.backward(dy.float())
    with config.use_precision("bf16" if dtype == torch.bfloat16 else "fp32"):
        tape = train.Tape(dtype, save=True)
        a = nhwc.Act(x, 16)
        a.rg = True
        ya = train.act_op(tape, a, act)
        tape.grads[id(ya)] = dy.clone()
        fn_b, _ = tape.ops[-1]       # the op's backward closure: leaves dx in the tape's gradient table under the input Act
        fn_b()
        got = tape.grads[id(a)]
    tol = 1e-6 if dtype == torch.float32 else 5e-3
    assert rel_err(ya.t.float(), yr.detach()) < tol
    assert rel_err(got.float(), xr.grad) < tol


def _golden(arch):
    return torch.load(os.path.join(GOLDEN, "backbone_%s.pt" % arch), weights_only=False)


_SPEC = {"mobilenet_v2": (O.mobilenet_v2_spec, O.mobilenet_v2_eca), "mobilenet_v3_small": (O.mobilenet_v3_small_spec, O.mobilenet_v3_small_eca),
         "mobilenet_v3_large": (O.mobilenet_v3_large_spec, O.mobilenet_v3_large_eca)}


@pytest.mark.parametrize("arch", ["mobilenet_v2", "mobilenet_v3_small", "mobilenet_v3_large"])
def test_mobilenet_backbones_fp32_vs_live_reference(arch):
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.backbone import get_backbone
    g = _golden(arch)
    spec_fn, fwd = _SPEC[arch]
    sd = O.seeded_state_dict(O.make_spec(spec_fn, 12, 2, 1), g["seed"])
    # yardstick for the gradients: the same step in fp64 on the CPU oracle (the live reference's own fp32 gradients are this far from it)
    leaf64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else (v.double() if v.is_floating_point() else v.clone()))
              for k, v in sd.items()}
    f64 = fwd(g["x"].double(), leaf64, "", True)
    (f64 * g["cot"].double()).sum().backward()
    n64 = {n: leaf64[n].grad.norm().item() for n in g["grads"]}
    with config.use_precision("fp32"):
        net = get_backbone(arch=arch, n_frames=4)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).eval()
        with torch.no_grad():
            fe = net(g["x"].to(dev)).cpu()
        net.train()
        for m in net.modules():                      # the golden was taken with dropout as identity
            if isinstance(m, torch.nn.Dropout):
                m.eval()
        ft = net(g["x"].to(dev))
        (ft * g["cot"].to(dev)).sum().backward()
    e_eval, e_train = rel_err(fe, g["feat_eval"]), rel_err(ft.detach().cpu(), g["feat_train"])
    # Parameters whose gradient is zero in exact arithmetic (the bias of a BatchNorm that feeds another BatchNorm through a linear
    # layer: the next normalisation removes it) carry only rounding noise (|g| ~ 1e-9 in fp32, 1e-17 in fp64): they are checked for
    # being that small, everything else relative to the fp64 gradient — by norm, like the golden records the reference's gradients.
    gmax = max(n64.values())
    live = [n for n in g["grads"] if n64[n] > 1e-6 * gmax]
    errs, tiny = {}, 0.0
    for name, p in net.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
        gn = p.grad.detach().double().norm().item()
        if name in live:
            errs[name] = abs(gn - n64[name]) / n64[name]
        else:
            tiny = max(tiny, gn / gmax)
    ref_live = sorted(abs(g["grads"][n]["norm"] - n64[n]) / n64[n] for n in live)
    mp, wp = ref_live[len(ref_live) // 2], ref_live[-1]
    vals = sorted(errs.values())
    med, worst = vals[len(vals) // 2], vals[-1]
    bn_err = max(rel_err(net.state_dict()[k].float().cpu(), v.float()) for k, v in g["bn"].items() if v.is_floating_point())
    print("\n[%s fp32] eval features %.2e, train features %.2e, BN statistics %.2e | %d gradient norms vs fp64: median %.2e worst %.2e (%s); "
          "the reference's own fp32 run vs fp64: median %.2e worst %.2e | %d exactly-zero gradients: largest %.1e of the largest norm"
          % (arch, e_eval, e_train, bn_err, len(live), med, worst, max(errs, key=errs.get), mp, wp, len(g["grads"]) - len(live), tiny))
    assert len(errs) == len(live) and len(live) > 0.8 * len(g["grads"])
    assert e_eval < 1e-4 and e_train < 1e-4 and bn_err < 1e-4
    assert tiny < 1e-5
    # at B = 2, 64x64 the last stages normalise over 8 values per channel: the reference's fp32 run is itself only this close to fp64
    assert med < max(10 * mp, 1e-4) and worst < max(10 * wp + 1e-3, 5e-2)


@pytest.mark.parametrize("arch", ["mobilenet_v2", "mobilenet_v3_small"])
def test_mobilenet_backbones_bf16_tensor_core_path(arch):
    """bf16: 1x1 convolutions on the tcgen05 kernel, depthwise / squeeze-excite / activations in bf16 storage. Eval features within
    the storage noise of ~50 layers of the fp32 reference, train step finite with every gradient present."""
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.backbone import get_backbone
    g = _golden(arch)
    spec_fn, fwd = _SPEC[arch]
    sd = O.seeded_state_dict(O.make_spec(spec_fn, 12, 2, 1), g["seed"])
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(8, 12, 128, 128, generator=gen)
    with torch.no_grad():
        ref = fwd(x, {k: v.clone() for k, v in sd.items()}, "", False)
    with config.use_precision("bf16"):
        net = get_backbone(arch=arch, n_frames=4)
        net.load_state_dict(sd, strict=True)
        net = net.to(dev).eval()
        with torch.no_grad():
            fe = net(x.to(dev)).cpu()
        net.train()
        ft = net(x.to(dev))
        ft.square().mean().backward()
    e = rel_err(fe, ref)
    print("\n[%s bf16] eval features vs fp32 oracle %.3e" % (arch, e))
    assert e < 5e-2     # ~50 stored bf16 tensors deep (2^-9 each, partly coherent): measured 2.6e-2 / 3.1e-2
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0 for p in net.parameters())
