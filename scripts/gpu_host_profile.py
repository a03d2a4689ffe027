"""cProfile of the host side of one K-expert training step (the step is host-bound when kernels are short)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf, loss as L, optim
from pmoe_b200.model.moe import get_model

K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
cfg = conf.stage2_model_cfg("moe", K)
model = get_model(cfg).cuda().train()
opt = optim.FusedAdam(model.parameters(), lr=2e-4, amsgrad=True)
g = torch.Generator().manual_seed(1)
images = torch.rand(B, 4, 3, 224, 224, generator=g).cuda()
speed = torch.rand(B, 1, generator=g).cuda()
command = torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=g), 6).float().cuda()
control = (torch.rand(B, 2, generator=g) * 2 - 1).cuda()
target = torch.rand(B, 1, generator=g).cuda()


def step():
    opt.zero_grad(set_to_none=True)
    d, sp = model(images, speed, command)
    L.moe_loss(d, sp, control, target.clone(), cfg.loss_coefs).backward()
    opt.step(max_grad_norm=1.0)


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(2):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
