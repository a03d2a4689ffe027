"""Who-waits-for-whom accounting of every tensor-core conv launch of one PU-Net inference step (tuning aid).
python scripts/gpu_conv_dbg.py [B] -> table of per-tile cycles by layer tag: MMA thread total, its waits for TMA data and
for a free accumulator, epilogue total and its waits."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pmoe_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = bench.build_punet().cuda().eval()
x = torch.rand(B, 4, 3, 224, 224, device="cuda")
with torch.no_grad():
    net(x)
torch.cuda.synchronize()
cnt = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
agg = collections.OrderedDict()
orig = ops.conv_tc


def wrapped(*a, **kw):
    cnt.zero_()
    _lib.lib().pmoe_conv_tc_set_debug(cnt.data_ptr())
    r = orig(*a, **kw)
    torch.cuda.synchronize()
    _lib.lib().pmoe_conv_tc_set_debug(None)
    c = cnt.view(148, 16).double()
    live = c[:, 7] > 0
    if live.any():
        c = c[live]
        d = agg.setdefault(kw.get("tag", "?"), [0, torch.zeros(16, dtype=torch.float64)])
        d[0] += 1
        d[1] += c.sum(0).cpu()
    return r


ops.conv_tc = wrapped
with torch.no_grad():
    net(x)
print("%-28s %5s %8s | %8s %8s %8s %8s | %8s %8s %8s | %6s %6s %6s %6s %6s" % ("tag", "n", "tiles", "prod_w", "mma_wTMA", "mma_wAcc", "mma_tot", "epi_wAcc", "epi_tot", "epi_wSt", "bar1", "ldtm", "work", "f+bar2", "issue"))
for tag, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1][3].item()):
    t = max(v[7].item(), 1.0)
    print("%-28s %5d %8d | %8.0f %8.0f %8.0f %8.0f | %8.0f %8.0f %8.0f | %6.0f %6.0f %6.0f %6.0f %6.0f" % (
        tag, n, t, v[0] / t, v[1] / t, v[2] / t, v[3] / t, v[4] / t, v[5] / t, v[6] / t, v[8] / t, v[9] / t, v[10] / t, v[11] / t, v[12] / t))
