"""Per-kernel-kind time breakdown of one training step (MoE, K experts) on the GPU. Not a bench: diagnosis only.
python scripts/gpu_train_profile.py [--k 6] [--batch 64] [--type moe] -> gpurun_out/train_profile.json"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf, loss as L, profiler
from pmoe_b200.model.moe import get_model

ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=6)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--type", default="moe")
ap.add_argument("--hw", type=int, default=224)
args = ap.parse_args()
torch.manual_seed(0)
cfg = conf.stage2_model_cfg(args.type, args.k)
model = get_model(cfg).cuda().train()
B = args.batch
g = torch.Generator().manual_seed(1234)
images = torch.rand(B, 4, 3, args.hw, args.hw, generator=g).cuda()
speed = (torch.rand(B, 1, generator=g) * 1.2).cuda()
command = torch.nn.functional.one_hot(torch.randint(0, 6, (B,), generator=g), 6).float().cuda()
control = (torch.rand(B, 2, generator=g) * 2 - 1).cuda()
target = torch.rand(B, 1, generator=g).cuda()


def step():
    for p in model.parameters():
        p.grad = None
    dist, sp = model(images, speed, command)
    loss = L.moe_loss(dist, sp, control, target.clone(), cfg.loss_coefs)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
profiler.reset()
profiler.enable_events(True)
step()
torch.cuda.synchronize()
recs = profiler.records()
profiler.enable_events(False)
by = {}
for kind, t, f, b, tag in recs:
    key = kind if kind not in ("conv_tc", "conv_simt", "conv_wgrad", "conv_wgrad_tc") else kind + (" wgrad" if tag.startswith("wgrad") else (" dgrad" if tag.startswith("dgrad") else ""))
    d = by.setdefault(key, {"ms": 0.0, "n": 0, "flops": 0.0, "bytes": 0.0})
    d["ms"] += t
    d["n"] += 1
    d["flops"] += f
    d["bytes"] += b
HBM_PEAK = 6551.0
try:
    HBM_PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
tot = sum(d["ms"] for d in by.values())
rows = sorted(by.items(), key=lambda kv: -kv[1]["ms"])
print("step %.2f ms (%.1f samples/s), loss %.4f; sum of kernel times %.2f ms over %d launches" % (ms, B / ms * 1e3, loss.item(), tot, len(recs)))
for k, d in rows:
    d["tflops"] = d["flops"] / max(d["ms"], 1e-9) / 1e9
    d["gbs"] = d["bytes"] / max(d["ms"], 1e-9) / 1e6   # algorithmic bytes (each operand once) / time
    d["hbm_frac"] = d["gbs"] / HBM_PEAK if d["bytes"] else None
    print("%-24s %8.3f ms %5.1f%% n=%4d %8.1f TF %8.0f GB/s %s" % (k, d["ms"], 100 * d["ms"] / tot, d["n"], d["tflops"], d["gbs"],
                                                               ("%.0f%% of HBM peak" % (100 * d["hbm_frac"])) if d["bytes"] else ""))
# slowest individual launches
top = sorted(recs, key=lambda r: -r[1])[:25]
for kind, t, f, b, tag in top:
    print("  %-16s %-44s %7.3f ms %8.1f TF %8.0f GB/s" % (kind, tag, t, f / max(t, 1e-9) / 1e9, b / max(t, 1e-9) / 1e6))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"ms_per_step": ms, "batch": B, "k": args.k, "hbm_peak_gbs": HBM_PEAK, "rows": rows, "top": top, "all": recs}, open("gpurun_out/train_profile_k%d_b%d.json" % (args.k, B), "w"), indent=1)
