"""B = 1 tick latency of the deployment path (pmoe_b200.agent.RealtimeSampler): eager launches vs one CUDA-graph replay,
host frame in -> host action out, for the 3-expert mixture (conf/stage_2.yaml default) and, with --pmoe, the full PMoE
(mixture + PU-Net expert: 10 U-Net passes per tick)."""
import argparse
import json
import os
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf
from pmoe_b200.agent import RealtimeSampler
from pmoe_b200.model.moe import get_model

ap = argparse.ArgumentParser()
ap.add_argument("--ticks", type=int, default=30)
ap.add_argument("--pmoe", action="store_true")
ap.add_argument("--out", default=None)
a = ap.parse_args()
torch.manual_seed(0)


def build(kind):
    if kind == "moe":
        return get_model(conf.stage2_model_cfg("moe", 3)).cuda().eval()
    import bench
    td = tempfile.mkdtemp()
    punet = bench.build_punet()
    torch.save({"unet": punet.unet.state_dict()}, os.path.join(td, "unet.pth"))
    torch.save({"model": punet.state_dict()}, os.path.join(td, "punet.pth"))
    cfg = conf.stage2_model_cfg("pmoe", 3)
    cfg.punet.model_path, cfg.punet_path = os.path.join(td, "unet.pth"), os.path.join(td, "punet.pth")
    torch.save(get_model(conf.stage2_model_cfg("moe", 3)).state_dict(), os.path.join(td, "moe.pth"))
    cfg.pmoe.moe_dir = os.path.join(td, "moe.pth")
    cfg.device = "cpu"
    return get_model(cfg).cuda().eval()


res = {}
hs, ws = 600, 800
g = torch.Generator().manual_seed(1)
frames = [torch.randint(0, 256, (hs, ws, 3), generator=g, dtype=torch.uint8).numpy() for _ in range(4)]
for kind in (["moe", "pmoe"] if a.pmoe else ["moe"]):
    model = build(kind)
    for mode in ("eager", "graph"):
        s = RealtimeSampler(model, (hs, ws), graph=(mode == "graph"))
        for i in range(5):
            s.step(frames[i % 4], 0.4, i % 6)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.ticks):
            s.step(frames[i % 4], 0.4, i % 6)
        ms = (time.perf_counter() - t0) / a.ticks * 1e3
        res["%s_%s_ms_per_tick" % (kind, mode)] = ms
        print("%-5s %-5s %.2f ms / tick (host frame in -> host action out)" % (kind, mode, ms), flush=True)
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
