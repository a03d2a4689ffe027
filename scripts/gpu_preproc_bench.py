"""Throughput of the GPU input pipeline (crop + Pillow-exact bilinear resize + ToTensor + stacking) on device-resident
decoded frames, against the HBM roofline. Algorithmic bytes per frame: the cropped source rows read once + the fp32 output
written once (the uint8 intermediate is not counted). Writes profiles/r01_preproc_bench.json when run with --out."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200.preproc import FramePreprocessor

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--hs", type=int, default=600)
ap.add_argument("--ws", type=int, default=800)
ap.add_argument("--out", default=None)
a = ap.parse_args()
pp = FramePreprocessor((125, 90), (224, 224))
x = torch.randint(0, 256, (a.frames, a.hs, a.ws, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(a.frames, 3, 224, 224, device="cuda")
for _ in range(3):
    pp(x, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    pp(x, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
rd = a.frames * (a.hs - 215) * a.ws * 3
wr = out.numel() * 4
peak = 6551.0
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
res = {"frames": a.frames, "src": [a.hs, a.ws], "ms": ms, "frames_per_s": a.frames / ms * 1e3, "algorithmic_gbps": (rd + wr) / ms / 1e6,
       "hbm_peak_gbps": peak, "frac": (rd + wr) / ms / 1e6 / peak, "launches": 2}
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
