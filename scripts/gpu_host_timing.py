"""Wall-clock split of the host side of one training step: forward, tape backward, autograd-engine remainder, optimizer."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf, loss as L, optim, train
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
T = {"tape_bwd": 0.0}
orig = train.TapeFunction.backward


def timed(ctx, *gouts):
    t0 = time.perf_counter()
    T["entry"] = t0
    r = orig(ctx, *gouts)
    T["exit"] = time.perf_counter()
    T["tape_bwd"] += T["exit"] - t0
    return r


train.TapeFunction.backward = staticmethod(timed)
acc = {"zero": 0.0, "fwd": 0.0, "loss": 0.0, "bwd_total": 0.0, "opt": 0.0}


def step():
    t = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    t1 = time.perf_counter(); acc["zero"] += t1 - t
    d, sp = model(images, speed, command)
    t2 = time.perf_counter(); acc["fwd"] += t2 - t1
    loss = L.moe_loss(d, sp, control, target.clone(), cfg.loss_coefs)
    t3 = time.perf_counter(); acc["loss"] += t3 - t2
    loss.backward()
    t4 = time.perf_counter(); acc["bwd_total"] += t4 - t3
    acc["pre"] = acc.get("pre", 0.0) + (T["entry"] - t3)
    acc["post"] = acc.get("post", 0.0) + (t4 - T["exit"])
    opt.step(max_grad_norm=1.0)
    acc["opt"] += time.perf_counter() - t4


for _ in range(3):
    step()
torch.cuda.synchronize()
for k in acc:
    acc[k] = 0.0
T["tape_bwd"] = 0.0
n = 3
t0 = time.perf_counter()
for _ in range(n):
    step()
th = time.perf_counter() - t0
torch.cuda.synchronize()
tt = time.perf_counter() - t0
print("K=%d B=%d: host %.1f ms/step, with GPU drain %.1f ms/step" % (K, B, th / n * 1e3, tt / n * 1e3))
for k, v in acc.items():
    print("  %-10s %.1f ms" % (k, v / n * 1e3))
print("  tape_bwd   %.1f ms (inside bwd_total; remainder = autograd engine: validation, AccumulateGrad)" % (T["tape_bwd"] / n * 1e3))
