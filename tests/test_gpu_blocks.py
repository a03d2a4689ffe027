"""Block-level GPU parity (fp32 mode, 1e-4) against the CPU oracle: conv3, ECA, EfficientConvBlock,
U-Net with inter_repr, PU-Net eval. Localises failures that the whole-model tests only detect."""
import os

import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import functional as O

pytestmark = pytest.mark.gpu


def _sd_grad(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("running_mean", "running_var")) else v.clone())
            for k, v in sd.items()}


def _compare_grads(module, sdg, tol):
    worst = (0.0, "")
    for name, p in module.named_parameters():
        og = sdg[name].grad
        if og is None:
            assert p.grad is None or p.grad.abs().sum() == 0, name
            continue
        assert p.grad is not None, name
        e = rel_err(p.grad.cpu(), og)
        if e > worst[0]:
            worst = (e, name)
    print("   worst grad err %.3e at %s" % worst)
    assert worst[0] < tol, worst


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("cin,cout", [(12, 64), (92, 3), (138, 64)])
def test_eca_conv_block_fp32(cin, cout, train):
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.basics import EfficientConvBlock
    spec = O.make_spec(O.eca_block_spec, cin, cout)
    sd = O.seeded_state_dict(spec, 3)
    x = torch.randn(3, cin, 24, 40, generator=torch.Generator().manual_seed(5))
    with config.use_precision("fp32"):
        blk = EfficientConvBlock(cin, cout)
        blk.load_state_dict(sd, strict=True)
        blk = blk.cuda().train(train)
        if train:
            y = blk(x.cuda())
            (y.cpu() * torch.linspace(-1, 1, y.numel()).reshape(y.shape)).sum().backward()
        else:
            with torch.no_grad():
                y = blk(x.cuda())
    sdg = _sd_grad(sd)
    ref = O.eca_conv_block(x, sdg, "", train)
    e = rel_err(y.detach().cpu(), ref.detach())
    print("\neca_conv_block %d->%d train=%s rel err %.3e" % (cin, cout, train, e))
    assert e < 1e-4
    if train:
        (ref * torch.linspace(-1, 1, ref.numel()).reshape(ref.shape)).sum().backward()
        _compare_grads(blk, sdg, 2e-3)


def test_eca_block_fp32_grad():
    from pmoe_b200 import config
    from pmoe_b200.model.blocks.basics import EfficientBlock
    x = torch.randn(2, 92, 16, 16, generator=torch.Generator().manual_seed(5))
    w = torch.randn(1, 1, 3, generator=torch.Generator().manual_seed(6))
    with config.use_precision("fp32"):
        m = EfficientBlock(92)
        m.load_state_dict({"conv.weight": w})
        m = m.cuda().train()
        xc = x.cuda()
        y = m(xc)
        (y.cpu() ** 2).sum().backward()
    wr = w.clone().requires_grad_(True)
    ref = O.eca(x, wr)
    (ref ** 2).sum().backward()
    print("\neca fwd %.3e  dw %.3e" % (rel_err(y.detach().cpu(), ref.detach()), rel_err(m.conv.weight.grad.cpu(), wr.grad)))
    assert rel_err(y.detach().cpu(), ref.detach()) < 1e-5
    assert rel_err(m.conv.weight.grad.cpu(), wr.grad) < 1e-4


def test_punet_eval_fp32_vs_golden(tmp_path):
    from pmoe_b200 import config
    from pmoe_b200.model.punet import PredictiveUnet
    g = torch.load(os.path.join(GOLDEN, "punet_stage1.pt"), weights_only=False)
    pc = dict(g["cfg"])
    sd = O.seeded_state_dict(O.make_spec(O.punet_spec, pc), g["seed"])
    ck = tmp_path / "unet.pth"
    torch.save({"unet": {k[len("unet."):]: v for k, v in sd.items() if k.startswith("unet.")}}, ck)
    pc["model_path"] = str(ck)
    with config.use_precision("fp32"):
        net = PredictiveUnet(**pc)
        net.load_state_dict(sd, strict=True)
        net = net.cuda().eval()
        with torch.no_grad():
            out = net(g["imgs"].cuda()).cpu()
    for f in range(out.shape[1]):
        print("\n frame %d rel err %.3e" % (f, rel_err(out[:, f, :, ::2, ::2], g["out_eval"][:, f])))
    assert rel_err(out[..., ::2, ::2], g["out_eval"]) < 1e-4
