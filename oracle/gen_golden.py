"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the LIVE reference
(/root/reference/PMoE, imported read-only with the shim of SURVEY.md App. D) on seeded weights and
inputs. Run in the build container only (the reference does not exist on the GPU box):

    python oracle/gen_golden.py

Weights are NOT stored: both sides rebuild them with oracle.functional.seeded_state_dict(spec, seed),
and `load_state_dict(strict=True)` into the reference module pins the key/shape contract.
"""
import copy
import json
import os
import sys
import tempfile
import types

import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import functional as O  # noqa: E402

REF = "/root/reference/PMoE"
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    thop = types.ModuleType("thop")
    thop.profile = lambda *a, **k: (0, 0)
    thop.clever_format = lambda x, f: ("0", "0")
    sys.modules["thop"] = thop
    sys.path[:0] = [REF, os.path.join(REF, "trainer")]


class AD(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def wrap(o):
    if isinstance(o, dict):
        return AD({k: wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [wrap(v) for v in o]
    return o


def synth_inputs(B, T, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, T, 3, H, W, generator=g)
    speed = torch.rand(B, 1, generator=g) * 1.2
    command = torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=g), 6).float()
    control = torch.rand(B, 2, generator=g) * 2 - 1
    target_speed = torch.rand(B, 1, generator=g)
    return images, speed, command, control, target_speed


def grad_summary(named_params, n_samples=8):
    """Per-parameter gradient fingerprint: L2 norm, sum, and values at fixed pseudo-random indices."""
    out = {}
    for name, prm in named_params:
        if prm.grad is None:
            continue
        gflat = prm.grad.detach().reshape(-1).double()
        gi = torch.Generator().manual_seed(len(name) * 7919 + gflat.numel())
        idx = torch.randint(0, gflat.numel(), (n_samples,), generator=gi)
        out[name] = {"norm": gflat.norm().item(), "sum": gflat.sum().item(), "idx": idx.tolist(),
                     "vals": gflat[idx].tolist()}
    return out


def bn_summary(sd):
    return {k: v.detach().clone() for k, v in sd.items() if k.endswith("running_mean") or k.endswith("running_var")
            or k.endswith("num_batches_tracked")}


def main():
    os.makedirs(OUT, exist_ok=True)
    import_reference()
    from model.blocks.unet import UNet
    from model.punet import PredictiveUnet
    from model.moe import get_model
    import loss as RL

    torch.manual_seed(0)
    torch.set_num_threads(8)
    specs_json = {}

    # ---------------------------------------------------------------- U-Net, stage-0 train step (config 1 shape, reduced H/W)
    spec = O.make_spec(O.unet_spec, 3, 23)
    sd = O.seeded_state_dict(spec, 11)
    net = UNet(in_features=3, out_features=23, gamma=2, b=1, inter_repr=False)
    net.load_state_dict(sd, strict=True)
    specs_json["unet"] = {k: list(v.shape) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    img = torch.rand(2, 3, 32, 32, generator=g)
    mask = torch.randint(0, 23, (2, 32, 32), generator=g)
    net.eval()
    with torch.no_grad():
        logits_eval = net(img)
    net.train()
    logits = net(img)
    lossv = RL.cross_entropy_tversky_weighted_loss(logits, mask)
    lossv.backward()
    torch.save({"img": img, "mask": mask, "logits_eval": logits_eval, "logits_train": logits.detach(), "loss": lossv.detach(),
                "dice_w": RL.class_dice(logits.detach(), mask), "tversky": RL.tversky_loss(logits.detach(), mask),
                "grads": grad_summary(net.named_parameters()), "bn": bn_summary(net.state_dict()), "seed": 11},
               os.path.join(OUT, "unet_stage0.pt"))
    print("unet_stage0 loss", lossv.item())

    # U-Net with inter_repr and odd-ish (non /16) size is not used by any conf; skip.

    # ---------------------------------------------------------------- PU-Net (stage 1): eval forward + train step with BPTT
    tmp = tempfile.mkdtemp()
    pc = dict(past_frames=4, future_frames=3, in_features=3, num_classes=23, gamma=2, b=1, inter_repr=False,
              unet_inter_repr=False, model_name="unet", model_path=os.path.join(tmp, "unet.pth"))
    torch.save({"unet": sd}, pc["model_path"])
    pspec = O.make_spec(O.punet_spec, pc)
    psd = O.seeded_state_dict(pspec, 12)
    pun = PredictiveUnet(**pc)
    pun.load_state_dict(psd, strict=True)
    specs_json["predictive_unet"] = {k: list(v.shape) for k, v in pun.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    # B=2 at 64x64: train-mode BN at the bottleneck then normalises over 32 values per channel. (B=1 at
    # 32x32 leaves 4 and makes the step chaotic: 1 vs 8 CPU threads already differ by 30% in some grads.)
    imgs = torch.rand(2, 4, 3, 64, 64, generator=g)
    masks = torch.randint(0, 23, (2, 3, 64, 64), generator=g)
    pun.eval()
    with torch.no_grad():
        out_eval = pun(imgs)[..., ::2, ::2].clone()  # stored subsampled to keep the fixture small
    pun.train()
    out_full = pun(imgs)
    out_tr = out_full[..., ::2, ::2]
    crit = RL.AutoregressiveCriterion(n_target_frames=3, loss_type="tversky")
    pl = crit(out_full, masks)
    pl.backward()
    torch.save({"imgs": imgs, "masks": masks, "out_eval": out_eval, "out_train": out_tr.detach().clone(), "loss": pl.detach(),
                "grads": grad_summary(pun.named_parameters()), "bn": bn_summary(pun.state_dict()), "seed": 12, "cfg": pc},
               os.path.join(OUT, "punet_stage1.pt"))
    print("punet_stage1 loss", pl.item())

    # ---------------------------------------------------------------- stage-2 models
    full = wrap(yaml.safe_load(open(os.path.join(REF, "conf", "stage_2.yaml")))).model
    full.backbone.rgb.pretrained = False
    full.device = "cpu"
    full.verbose = False
    for k in ("action_head", "speed_encoder", "command_encoder", "speed_prediction"):
        full[k].dropout = 0.0  # parity runs disable dropout (SURVEY §7 hard part 5)
    full.punet.future_frames = 3  # keep the fixture small; the constructor path is identical
    full.punet.model_path = pc["model_path"]
    full.punet.model_name = "unet"
    punet_ckpt = os.path.join(tmp, "punet.pth")
    torch.save({"model": psd}, punet_ckpt)
    full.punet_path = punet_ckpt

    def plain(o):
        if isinstance(o, dict):
            return {k: plain(v) for k, v in o.items()}
        if isinstance(o, list):
            return [plain(v) for v in o]
        return o

    B, H, W = 2, 64, 64
    images, speed, command, control, target_speed = synth_inputs(B, 4, H, W)
    for mtype, seed in (("moe", 21), ("moe_alt", 22), ("moe_shared", 23)):
        cfg = copy.deepcopy(full)
        cfg.type = mtype
        ocfg = plain(cfg)
        spec_fn = O.moe_shared_spec if mtype == "moe_shared" else O.moe_spec
        msd = O.seeded_state_dict(O.make_spec(spec_fn, ocfg), seed)
        model = get_model(cfg)
        model.load_state_dict(msd, strict=True)
        specs_json[mtype] = {k: list(v.shape) for k, v in model.state_dict().items()}
        model.eval()
        with torch.no_grad():
            dist_e, sp_e = model(images, speed, command)
            torch.manual_seed(77)
            samp = model.sample(images, speed, command)
        model.train()
        dist, sp = model(images, speed, command)
        ts = target_speed.clone()
        lossv = RL.moe_loss(dist, sp, control, ts, cfg.loss_coefs)
        lossv.backward()
        rec = {"images": images, "speed": speed, "command": command, "control": control, "target_speed": target_speed,
               "probs_eval": dist_e.mixture_distribution.probs, "mean_eval": dist_e.component_distribution.base_dist.loc,
               "std_eval": dist_e.component_distribution.base_dist.scale, "speed_eval": sp_e, "sample_eval_seed77": samp,
               "probs_train": dist.mixture_distribution.probs.detach(), "mean_train": dist.component_distribution.base_dist.loc.detach(),
               "std_train": dist.component_distribution.base_dist.scale.detach(), "speed_train": sp.detach(),
               "log_prob_train": dist.log_prob(control).detach(), "loss": lossv.detach(),
               "target_speed_after": ts, "grads": grad_summary(model.named_parameters()),
               "bn": bn_summary(model.state_dict()), "seed": seed, "cfg": ocfg}
        torch.save(rec, os.path.join(OUT, "%s_stage2.pt" % mtype))
        print(mtype, "loss", lossv.item(), "argmax", dist_e.mixture_distribution.probs.argmax(1).tolist())
        if mtype == "moe":
            moe_sd_path = os.path.join(tmp, "moe.pth")
            torch.save(msd, moe_sd_path)
            full.pmoe.moe_dir = moe_sd_path

    B, H, W = 2, 64, 64
    images, speed, command, control, target_speed = synth_inputs(B, 4, H, W)
    for mtype, seed in (("punet", 31), ("punet_inter", 32)):
        cfg = copy.deepcopy(full)
        cfg.type = mtype
        ocfg = plain(cfg)
        ocfg["punet"]["inter_repr"] = (mtype == "punet_inter")
        esd = O.seeded_state_dict(O.make_spec(O.punet_expert_spec, ocfg), seed)
        model = get_model(cfg)
        model.load_state_dict(esd, strict=True)
        specs_json[mtype] = {k: list(v.shape) for k, v in model.state_dict().items()}
        model.eval()
        with torch.no_grad():
            a_e, s_e = model(images, speed, command)
        rec = {"images": images, "speed": speed, "command": command, "control": control, "target_speed": target_speed,
               "actions_eval": a_e, "speed_eval": s_e, "seed": seed, "cfg": ocfg,
               "requires_grad": {n: p.requires_grad for n, p in model.named_parameters()}}
        if mtype == "punet":
            model.train()
            a, s = model(images, speed, command)
            lossv = RL.punet_loss(a, s, control, target_speed, cfg.loss_coefs)
            lossv.backward()
            rec.update({"actions_train": a.detach(), "speed_train": s.detach(), "loss": lossv.detach(),
                        "grads": grad_summary(model.named_parameters())})
            print(mtype, "loss", lossv.item())
        torch.save(rec, os.path.join(OUT, "%s_stage2.pt" % mtype))

    cfg = copy.deepcopy(full)
    cfg.type = "pmoe"
    ocfg = plain(cfg)
    ocfg["punet"]["inter_repr"] = False
    fsd = O.seeded_state_dict(O.make_spec(O.pmoe_spec, ocfg), 41)
    model = get_model(cfg)
    model.load_state_dict(fsd, strict=True)
    specs_json["pmoe"] = {k: list(v.shape) for k, v in model.state_dict().items()}
    model.eval()
    with torch.no_grad():
        torch.manual_seed(99)
        a_e, dummy = model(images, speed, command)
    lossv = RL.pmoe_loss(a_e, dummy, control, target_speed, cfg.loss_coefs)
    torch.save({"images": images, "speed": speed, "command": command, "control": control, "actions_eval_seed99": a_e,
                "speed_dummy": dummy, "loss": lossv, "seed": 41, "cfg": ocfg,
                "requires_grad": {n: p.requires_grad for n, p in model.named_parameters()}},
               os.path.join(OUT, "pmoe_stage2.pt"))
    print("pmoe loss", lossv.item())

    json.dump(specs_json, open(os.path.join(OUT, "state_dict_specs.json"), "w"))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
