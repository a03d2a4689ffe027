"""How exactly does the tcgen05 path reproduce an fp32-accumulated convolution of the same bf16 operands? One 3x3 conv (64 -> 128,
B=8, 32x32): fraction of bf16 outputs that differ from round_bf16(fp32 reference) for the tensor-core kernel and for the CUDA-core
kernel (fp32 FMA chain), and the size of the differences in bf16 ulps."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import config, infer, nhwc
from pmoe_b200.nhwc import Act

dev = "cuda"
g = torch.Generator().manual_seed(0)
for (cin, cout, hw) in ((64, 128, 32), (128, 128, 56), (512, 512, 14)):
    x = torch.randn(8, cin, hw, hw, generator=g).abs().to(torch.bfloat16)
    conv = torch.nn.Conv2d(cin, cout, 3, 1, 1, bias=False)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
    ref64 = F.conv2d(x.double(), conv.weight.double(), None, 1, 1)
    ref = ref64.float().to(torch.bfloat16)                      # correctly rounded result
    conv = conv.to(dev)
    xa = Act(x.permute(0, 2, 3, 1).contiguous().to(dev), cin)
    for simt in (False, True):
        config.FORCE_SIMT = simt
        try:
            with config.use_precision("bf16"):
                y = infer.conv_eval([xa], conv).t[..., :cout].permute(0, 3, 1, 2).cpu()
        finally:
            config.FORCE_SIMT = False
        diff = (y.float() - ref.float()).abs()
        ulp = ref.float().abs().clamp_min(1e-30) * 2.0 ** -7
        frac = (y != ref).float().mean().item()
        rel = ((y.double() - ref64).norm() / ref64.norm()).item()
        rel_ref = ((ref.double() - ref64).norm() / ref64.norm()).item()
        print("%d->%d @%d %-10s: %.4f%% of the bf16 outputs differ from the correctly rounded fp64 result (max %.2f ulp); rel err vs fp64 %.3e "
              "(correct rounding alone: %.3e)" % (cin, cout, hw, "CUDA cores" if simt else "tcgen05", 100 * frac, (diff / ulp).max().item(), rel, rel_ref))
