"""Checkpoint interchange: a stage-0 checkpoint written in the reference trainer's layout
(train_0.py:313-338) from this package's modules round-trips, and — when the reference tree is present (this container) —
loads into the LIVE reference modules with strict=True and back."""
import os
import sys

import pytest
import torch

REF = "/root/reference/PMoE"


def test_stage0_checkpoint_round_trip(tmp_path):
    from pmoe_b200.model.blocks.unet import UNet
    from pmoe_b200.utils.io import save_checkpoint, load_checkpoint
    torch.manual_seed(0)
    net = UNet(3, 23)
    swa = torch.optim.swa_utils.AveragedModel(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, amsgrad=True)
    ck = {"epoch": 3, "iteration": 77, "unet": net.state_dict(), "unet-swa": swa.state_dict(), "optimizer": opt.state_dict(),
          "best": 0.5, "dice": 0.4, "e_loss": [1.0]}
    path = save_checkpoint(ck, True, str(tmp_path / "ckpt"), "unet-e3")
    assert os.path.exists(path) and os.path.exists(str(tmp_path / "ckpt" / "unet-best.pth"))
    back = load_checkpoint(path, "cpu")
    net2 = UNet(3, 23)
    net2.load_state_dict(back["unet"], strict=True)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    swa2 = torch.optim.swa_utils.AveragedModel(UNet(3, 23))
    swa2.load_state_dict(back["unet-swa"], strict=True)
    with pytest.raises(FileNotFoundError):
        load_checkpoint(str(tmp_path / "missing.pth"), "cpu")


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_state_dicts_interchange_with_the_live_reference():
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.gen_golden import import_reference
    import_reference()
    from model.blocks.unet import UNet as RefUNet                      # reference modules
    from model.blocks.backbone import get_backbone as ref_backbone
    from pmoe_b200.model.blocks.unet import UNet
    from pmoe_b200.model.blocks.backbone import get_backbone
    torch.manual_seed(1)
    mine, ref = UNet(3, 23), RefUNet(3, 23)
    ref.load_state_dict(mine.state_dict(), strict=True)
    mine.load_state_dict(ref.state_dict(), strict=True)
    for arch in ("resnet18", "resnet34", "resnet50"):
        a = get_backbone(arch=arch, n_frames=4, pretrained=False, gamma=2, b=1, n_channels=3)
        b = ref_backbone(arch=arch, n_frames=4, pretrained=False, gamma=2, b=1, n_channels=3)
        b.load_state_dict(a.state_dict(), strict=True)
        a.load_state_dict(b.state_dict(), strict=True)
