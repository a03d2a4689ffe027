"""Where does the bf16 training path lose gradient accuracy? The same step in fp32 (CUDA cores), bf16 on the CUDA-core
kernels and bf16 on the tensor-core kernels, each against the fp32 CPU oracle on bf16-rounded weights/inputs, for nets of
growing depth; plus run-to-run determinism of the forward. Prints median / worst normwise gradient error per setting."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import functional as O
from pmoe_b200 import config, train
from pmoe_b200.model.blocks.basics import conv3
from pmoe_b200.model.blocks.unet import UNet
from pmoe_b200.model.blocks.backbone import get_backbone

dev = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def rnd(sd):
    return {k: (v.to(torch.bfloat16).float() if (v.is_floating_point() and v.dim() >= 2) else v.clone()) for k, v in sd.items()}


def leaves(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}


def settings():
    yield "fp32", "fp32", False
    yield "bf16-simt", "bf16", True
    yield "bf16-tc", "bf16", False


def report(tag, make_module, sd, x, oracle_fn, up_seed=9, flags=None):
    sdg = leaves(sd)
    ref = oracle_fn(x, sdg)
    up = torch.randn(ref.shape, generator=torch.Generator().manual_seed(up_seed)) * 1e-2
    ref.backward(gradient=up)
    for name, prec, simt in settings():
        config.FORCE_SIMT = simt
        try:
            with config.use_precision(prec):
                m = make_module()
                m.load_state_dict(sd, strict=True)
                m = m.to(dev).train()
                y = m(x.to(dev))
                y.backward(gradient=up.to(dev))
        finally:
            config.FORCE_SIMT = False
        errs = {n: rel(p.grad, sdg[n].grad) for n, p in m.named_parameters() if sdg[n].grad is not None and p.grad is not None}
        vals = sorted(errs.values())
        worst = max(errs, key=errs.get)
        print("%-28s %-10s out %.2e | grads median %.2e p90 %.2e worst %.2e (%s)" % (tag, name, rel(y.detach(), ref.detach()), vals[len(vals) // 2],
                                                                                   vals[int(0.9 * (len(vals) - 1))], errs[worst], worst), flush=True)
        if name == "bf16-tc" and flags:
            for fl in flags:
                old = getattr(train, fl)
                setattr(train, fl, False)
                try:
                    with config.use_precision(prec):
                        m = make_module()
                        m.load_state_dict(sd, strict=True)
                        m = m.to(dev).train()
                        y = m(x.to(dev))
                        y.backward(gradient=up.to(dev))
                finally:
                    setattr(train, fl, old)
                errs = {n: rel(p.grad, sdg[n].grad) for n, p in m.named_parameters() if sdg[n].grad is not None and p.grad is not None}
                vals = sorted(errs.values())
                print("%-28s %-10s out %.2e | grads median %.2e worst %.2e   [%s off]" % (tag, name, rel(y.detach(), ref.detach()), vals[len(vals) // 2],
                                                                                        vals[-1], fl), flush=True)


g = torch.Generator().manual_seed(8)
# 1. one conv3 block
spec = O.make_spec(lambda sp, p: O.conv3_spec(sp, p, 64, 128))
sd = rnd(O.seeded_state_dict(spec, 2))
x = torch.randn(8, 64, 32, 32, generator=g).to(torch.bfloat16).float()
report("conv3 64->128 B8 32x32", lambda: conv3(64, 128), sd, x, lambda x_, s_: O.conv3_block(x_, s_, "", True))
# 2. U-Net at two sizes
for (B, H) in ((4, 64), (8, 128)):
    sd = rnd(O.seeded_state_dict(O.make_spec(O.unet_spec, 3, 23), 11))
    x = torch.rand(B, 3, H, H, generator=g).to(torch.bfloat16).float()
    report("unet B%d %dx%d" % (B, H, H), lambda: UNet(3, 23), sd, x, lambda x_, s_: O.unet(x_, s_, "", True))
# 3. ResNet18-ECA backbone
for (B, H) in ((8, 64), (8, 128)):
    sd = rnd(O.seeded_state_dict(O.make_spec(O.resnet18_spec, 12), 12))
    x = torch.rand(B, 12, H, H, generator=g).to(torch.bfloat16).float()
    report("resnet18-eca B%d %dx%d" % (B, H, H), lambda: get_backbone("resnet18", 4), sd, x, lambda x_, s_: O.resnet18_eca(x_, s_, "", True),
           flags=("FUSE_BN_CHAIN_SUMS", "FUSE_STRIDE2_DGRAD"))

# 4. determinism of the forward (same module, same input, three passes)
with config.use_precision("bf16"):
    m = get_backbone("resnet18", 4).to(dev).train()
    x = torch.rand(8, 12, 64, 64, generator=g).to(dev)
    with torch.no_grad():
        outs = [m(x).clone() for _ in range(3)]
    print("forward determinism (bf16 resnet18-eca): max |diff| run1-run0 %.3e, run2-run0 %.3e" % ((outs[1] - outs[0]).abs().max().item(),
                                                                                            (outs[2] - outs[0]).abs().max().item()))
