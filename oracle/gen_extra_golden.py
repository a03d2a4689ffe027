"""TEST INFRASTRUCTURE — goldens for the low-traffic rows of SURVEY.md §8(a) from the LIVE reference:
  a6   UNetECA (PMoE/model/blocks/unet.py:98-185): eval forward, train forward + backward, with and without inter_repr
  a20  dice_score, l1_gdl (PMoE/trainer/loss.py:20-31,58-83)
  a19  AutoregressiveCriterion with loss_type 'l1' / 'l2' (loss.py:86-118), class_dice / tversky_loss on their own
Writes tests/golden/unet_eca.pt and tests/golden/seg_losses_extra.pt. Run in the build container only:
    python oracle/gen_extra_golden.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import functional as O  # noqa: E402
from oracle.gen_golden import import_reference, grad_summary, bn_summary  # noqa: E402


def main():
    import_reference()
    from model.blocks.unet import UNetECA
    import loss as RL
    torch.set_num_threads(8)

    # ------------------------------------------------------------------ UNetECA
    spec = O.make_spec(O.unet_eca_spec, 3, 23)
    sd = O.seeded_state_dict(spec, 41)
    g = torch.Generator().manual_seed(4100)
    x = torch.rand(2, 3, 32, 48, generator=g)
    cot = torch.randn(2, 23, 32, 48, generator=g) * 1e-2
    cot_i = torch.randn(2, 512, generator=g) * 1e-2
    out = {"seed": 41, "x": x, "cot": cot, "cot_inter": cot_i}
    for inter in (False, True):
        net = UNetECA(in_features=3, out_features=23, gamma=2, b=1, inter_repr=inter)
        net.load_state_dict(sd, strict=True)
        out["keys"] = {k: list(v.shape) for k, v in net.state_dict().items()}
        net.eval()
        with torch.no_grad():
            r = net(x)
        ev = r if not inter else r[1]
        net.train()
        r = net(x)
        if inter:
            (r[1] * cot).sum().backward(retain_graph=True)
            (r[0] * cot_i).sum().backward()
            tr, tr_i = r[1].detach(), r[0].detach()
        else:
            (r * cot).sum().backward()
            tr, tr_i = r.detach(), None
        tag = "inter" if inter else "plain"
        out[tag] = {"logits_eval": ev, "logits_train": tr, "inter_train": tr_i, "grads": grad_summary(net.named_parameters()),
                    "bn": bn_summary(net.state_dict())}
        print("UNetECA", tag, tuple(tr.shape), "params", sum(p.numel() for p in net.parameters()))
    torch.save(out, os.path.join(ROOT, "tests", "golden", "unet_eca.pt"))

    # ------------------------------------------------------------------ losses
    g = torch.Generator().manual_seed(4200)
    B, T, Cc, H, W = 2, 3, 23, 20, 28
    inputs = torch.randn(B, T, Cc, H, W, generator=g) * 1.5
    targets = torch.randint(0, Cc, (B, T, H, W), generator=g)
    # make the argmax agree with the target on about half of the pixels so that the dice counts are not trivial
    agree = torch.rand(B, T, H, W, generator=g) < 0.5
    boost = torch.zeros_like(inputs).scatter_(2, targets.unsqueeze(2), 8.0) * agree.unsqueeze(2)
    inputs = inputs + boost
    rec = {"inputs": inputs, "targets": targets}
    rec["dice_score"] = RL.dice_score(inputs[:, -1], targets[:, -1])
    rec["class_dice"] = RL.class_dice(inputs[:, -1], targets[:, -1])
    rec["tversky"] = RL.tversky_loss(inputs[:, -1], targets[:, -1])
    for name, fn in (("l1_gdl", lambda a, t: RL.l1_gdl(a, t)),
                     ("ar_l1", lambda a, t: RL.AutoregressiveCriterion(T, "l1")(a, t)),
                     ("ar_l2", lambda a, t: RL.AutoregressiveCriterion(T, "l2")(a, t)),
                     ("ar_tversky", lambda a, t: RL.AutoregressiveCriterion(T, "tversky")(a, t))):
        leaf = inputs.clone().requires_grad_(True)
        val = fn(leaf, targets)
        val.backward()
        rec[name] = val.detach()
        rec[name + "_grad"] = leaf.grad.clone()
        print(name, val.item(), leaf.grad.abs().sum().item())
    torch.save(rec, os.path.join(ROOT, "tests", "golden", "seg_losses_extra.pt"))


if __name__ == "__main__":
    main()
