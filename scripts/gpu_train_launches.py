"""Every GPU kernel one training micro-step launches (forward + moe_loss + backward of the K-expert mixture), grouped by
kernel name with launch counts and device time (torch.profiler / CUPTI, eager tape). Separates this library's kernels from
the ATen glue (fills, copies, small elementwise ops) so that the fixed per-step launch overhead — what limits strong scaling
at 64 samples per GPU — can be tracked.   usage: gpu_train_launches.py [K] [B] [out.json]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf, loss as L, optim
from pmoe_b200.model.moe import get_model

K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out = sys.argv[3] if len(sys.argv) > 3 else None
dev = "cuda"
torch.manual_seed(0)
cfg = conf.stage2_model_cfg("moe", K)
model = get_model(cfg).to(dev).train()
opt = optim.FusedAdam(model.parameters(), lr=2e-4, amsgrad=True)
g = torch.Generator().manual_seed(1)
d = {"images": torch.rand(B, 4, 3, 224, 224, generator=g).to(dev), "speed": (torch.rand(B, 1, generator=g) * 1.2).to(dev),
     "command": torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=g), 6).float().to(dev),
     "control": (torch.rand(B, 2, generator=g) * 2 - 1).to(dev), "target": torch.rand(B, 1, generator=g).to(dev)}
torch.distributions.Distribution.set_default_validate_args(False)


def step():
    opt.zero_grad(set_to_none=True)
    dist_, sp = model(d["images"], d["speed"], d["command"])
    L.moe_loss(dist_, sp, d["control"], d["target"].clone(), cfg.loss_coefs).backward()
    opt.step(max_grad_norm=1.0)


for _ in range(3):
    step()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    r = rows.setdefault(ev.name, [0, 0.0])
    r[0] += 1
    r[1] += ev.device_time
ours = ("conv_tc", "conv_wgrad", "conv_simt", "bn_", "affine", "maxpool", "eca_", "scale_channels", "channel_s", "prod_channel", "axpy",
        "nchw", "nhwc", "gate_", "moe_loss", "dropout", "mt_", "pack_", "unpack_", "cvt_f64", "act_reduce", "l1_mse", "segloss", "onehot")
tot_n = sum(r[0] for r in rows.values())
tot_us = sum(r[1] for r in rows.values())
mine_n = sum(r[0] for k, r in rows.items() if any(o in k for o in ours))
mine_us = sum(r[1] for k, r in rows.items() if any(o in k for o in ours))
print("K=%d B=%d: %d kernel launches, %.2f ms of device time; this library: %d launches / %.2f ms; ATen + memset/memcpy glue: %d launches / %.2f ms"
      % (K, B, tot_n, tot_us / 1e3, mine_n, mine_us / 1e3, tot_n - mine_n, (tot_us - mine_us) / 1e3))
top = sorted(rows.items(), key=lambda kv: -kv[1][0])
for name, (n, us) in top[:45]:
    print("%6d  %9.1f us  %s" % (n, us, name[:150]))
if out:
    json.dump({"K": K, "B": B, "launches": tot_n, "device_ms": tot_us / 1e3, "library_launches": mine_n, "library_ms": mine_us / 1e3,
               "kernels": {k: {"n": v[0], "us": v[1]} for k, v in top}}, open(out, "w"), indent=1)
