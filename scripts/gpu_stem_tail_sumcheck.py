"""Scratch diagnostic: where does sum(dx) of the fused stem-tail backward come from? Compares the reduce kernel's s1/s2 with the
sums implied by the apply kernel's dx and with a torch evaluation of the same routing."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200._lib import lib, check, view4, stream_ptr

dev = "cuda"
n, h, w, c = 16, 224, 224, 64
g = torch.Generator(device=dev).manual_seed(5)
x = (torch.randn(n, h, w, c, generator=g, device=dev) * 1.5 + 0.3).to(torch.bfloat16)
gamma = torch.rand(c, generator=g, device=dev) + 0.5
beta = torch.randn(c, generator=g, device=dev) * 0.3
xd = x.double()
N = n * h * w
mean64 = xd.mean(dim=(0, 1, 2))
rstd64 = 1.0 / torch.sqrt(xd.var(dim=(0, 1, 2), unbiased=False) + 1e-5)
scale = (gamma.double() * rstd64).float()
shift = (beta.double() - mean64 * gamma.double() * rstd64).float()
mean, rstd = mean64.float(), rstd64.float()
p = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=dev)
xm = torch.empty_like(p)
idx = torch.empty(p.shape, dtype=torch.uint8, device=dev)
vx, vp = view4(x), view4(p)
check(lib().pmoe_bn_relu_maxpool_fwd(C.byref(vx), scale.data_ptr(), shift.data_ptr(), C.byref(vp), idx.data_ptr(), xm.data_ptr(), stream_ptr()), "fwd")
dp = torch.randn(p.shape, generator=g, device=dev).to(torch.bfloat16)
s1 = torch.zeros(c, dtype=torch.float64, device=dev)
s2 = torch.zeros(c, dtype=torch.float64, device=dev)
vdp = view4(dp)
check(lib().pmoe_bn_relu_maxpool_bwd_reduce(C.byref(vdp), xm.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                            s1.data_ptr(), s2.data_ptr(), None, stream_ptr()), "reduce")
dx = torch.empty_like(x)
vdx = view4(dx)
check(lib().pmoe_bn_relu_maxpool_bwd_apply(C.byref(vdp), idx.data_ptr(), C.byref(vx), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                           rstd.data_ptr(), gamma.data_ptr(), s1.data_ptr(), s2.data_ptr(), 1.0 / N, C.byref(vdx), None, None,
                                           None, stream_ptr()), "apply")
# torch routing on the same operands
z = torch.relu(torch.addcmul(shift, x.float(), scale)).to(torch.bfloat16).float().permute(0, 3, 1, 2).requires_grad_(True)
pp = torch.nn.functional.max_pool2d(z, 3, 2, 1)
pp.backward(dp.float().permute(0, 3, 1, 2))
dz_t = (z.grad * (z > 0)).permute(0, 2, 3, 1).double()          # routed and masked
xh = (xd - mean64) * rstd64
print("s1 kernel vs torch       :", ((s1 - dz_t.sum(dim=(0, 1, 2))).abs().max() / dz_t.abs().sum(dim=(0, 1, 2)).max()).item())
print("s2 kernel vs torch       :", ((s2 - (dz_t * xh).sum(dim=(0, 1, 2))).abs().max() / dz_t.abs().sum(dim=(0, 1, 2)).max()).item())
A = gamma.double() * rstd64
ref = A * (dz_t - s1 / N - xh * s2 / N)
print("dx vs formula (norm)     :", ((dx.double() - ref).norm() / ref.norm()).item())
err = dx.double() - ref
print("sum err / sum|ref|       :", (err.sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))).max().item())
print("sum ref / sum|ref|       :", (ref.sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))).max().item())
print("sum dx / sum|dx|         :", (dx.double().sum(dim=(0, 1, 2)).abs() / dx.double().abs().sum(dim=(0, 1, 2))).max().item())
refb = ref.to(torch.bfloat16).double()
print("sum bf16(ref) / sum|ref| :", (refb.sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))).max().item())
on = dz_t != 0
print("fraction routed          :", on.double().mean().item())
e_on = (err * on).sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))
e_off = (err * (~on)).sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))
print("sum err on routed pixels :", e_on.max().item(), " on the others:", e_off.max().item())
rb = (refb - ref)
print("bf16(ref)-ref on routed  :", ((rb * on).sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))).max().item(), " others:",
      ((rb * (~on)).sum(dim=(0, 1, 2)).abs() / ref.abs().sum(dim=(0, 1, 2))).max().item())
