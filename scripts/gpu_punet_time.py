"""Quick timing of PU-Net inference (config 2 shape) on one B200."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import functional as O
from pmoe_b200.model.punet import PredictiveUnet
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pc = dict(past_frames=4, future_frames=6, in_features=3, num_classes=23, gamma=2, b=1, inter_repr=False, unet_inter_repr=False,
          model_name="unet", model_path="/tmp/unet.pth")
sd = O.seeded_state_dict(O.make_spec(O.punet_spec, pc), 1)
torch.save({"unet": {k[5:]: v for k, v in sd.items() if k.startswith("unet.")}}, pc["model_path"])
net = PredictiveUnet(**pc); net.load_state_dict(sd); net = net.cuda().eval()
x = torch.rand(B, 4, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y = net(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"B": B, "ms": ms, "samples_per_s": B / ms * 1e3, "tflops": 730.82e9 * B / ms / 1e9, "mem_gb": torch.cuda.max_memory_allocated() / 2**30}))
