"""One full training step (forward + moe_loss + backward + fused clip/Adam) of the K-expert mixture between cudaProfilerStart /
cudaProfilerStop, for `ncu --profile-from-start off` captures of the training-side kernels (weight gradient, gating / loss,
optimizer, pooling, BatchNorm reductions ...).   python scripts/gpu_ncu_train_step.py [K] [B] [preproc]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import conf, loss as L, optim
from pmoe_b200.model.moe import get_model

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
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


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
if len(sys.argv) > 3:   # the GPU input pipeline in the same capture
    from pmoe_b200 import preproc
    frames = torch.randint(0, 256, (64, 600, 800, 3), dtype=torch.uint8, generator=g).to(dev)
    preproc.FramePreprocessor()(frames)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu train step ok")
