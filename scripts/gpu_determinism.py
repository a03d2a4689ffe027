"""Run the same fp32 U-Net training step repeatedly from identical state and report run-to-run gradient differences
(summation-order noise is ~1e-6; anything near 1e-3 points at a race or an uninitialised read)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import config
from pmoe_b200.model.blocks.unet import UNet

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
config.set_precision(prec)
torch.manual_seed(0)
net = UNet(3, 23).cuda().train()
init = {k: v.clone() for k, v in net.state_dict().items()}
g = torch.Generator().manual_seed(1)
img = torch.rand(2, 3, 32, 32, generator=g).cuda()
up = torch.randn(2, 23, 32, 32, generator=g).cuda() * 1e-3
ref = None
for it in range(16):
    net.load_state_dict(init)
    for p in net.parameters():
        p.grad = None
    out = net(img)
    out.backward(gradient=up)
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    if ref is None:
        ref, ref_out = grads, out.detach().clone()
        continue
    worst = max((((grads[n] - ref[n]).norm() / (ref[n].norm() + 1e-30)).item(), n) for n in grads)
    eo = ((out.detach() - ref_out).norm() / ref_out.norm()).item()
    print("run %2d: out diff %.2e | worst grad diff %.2e (%s)" % (it, eo, worst[0], worst[1]), flush=True)
