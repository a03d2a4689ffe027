"""Per-launch CUDA-event profile of one PU-Net inference step (tag, ms, TFLOP/s). Writes gpurun_out/layer_profile.json."""
import json, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pmoe_b200 import profiler
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = bench.build_punet().cuda().eval()
x = torch.rand(B, 4, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(2): net(x)
    torch.cuda.synchronize()
    profiler.enable_events(True); net(x); recs = profiler.records(); profiler.enable_events(False)
agg = collections.OrderedDict()
for kind, ms, fl, by, tag in recs:
    d = agg.setdefault((kind, tag), [0.0, 0.0, 0])
    d[0] += ms; d[1] += fl; d[2] += 1
tot = sum(v[0] for v in agg.values())
rows = sorted(((k, v) for k, v in agg.items()), key=lambda kv: -kv[1][0])
out = [{"kind": k[0], "tag": k[1], "ms": v[0], "share": v[0] / tot, "launches": v[2], "tflops": (v[1] / v[0] / 1e9) if v[0] > 0 else 0} for k, v in rows]
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"B": B, "total_ms": tot, "rows": out}, open("gpurun_out/layer_profile.json", "w"), indent=1)
print("total ms", tot)
for r in out[:32]:
    print("%-14s %-28s %8.3f ms %5.1f%% n=%3d %7.1f TF" % (r["kind"], r["tag"], r["ms"], 100 * r["share"], r["launches"], r["tflops"]))
