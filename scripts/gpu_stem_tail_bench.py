"""The three launches of csrc/stem_tail.cu (bn1 -> ReLU -> MaxPool2d(3,2,1) of the ResNet stem, training) alone at the bench
geometry: CUDA-event time, algorithmic bytes (every operand read / written once) and GB/s against the measured HBM peak.
Also the ncu target for these kernels.   python scripts/gpu_stem_tail_bench.py [B] [out.json]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import train
from pmoe_b200._lib import lib, check, view4, stream_ptr

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
out = sys.argv[2] if len(sys.argv) > 2 else None
dev, H, W, Cc = "cuda", 224, 224, 64
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(B, H, W, Cc, generator=g) * 1.5 + 0.3).to(torch.bfloat16).to(dev)   # a ReLU output, as in the stem
gamma = (torch.rand(Cc, generator=g) + 0.5).to(dev)
beta = (torch.randn(Cc, generator=g) * 0.3).to(dev)
dp = torch.randn(B, H // 2, W // 2, Cc, generator=g).to(torch.bfloat16).to(dev)
xd = x.double()
mean = xd.mean(dim=(0, 1, 2))
rstd = 1.0 / torch.sqrt(xd.var(dim=(0, 1, 2), unbiased=False) + 1e-5)
scale = (gamma.double() * rstd).float()
shift = (beta.double() - mean * gamma.double() * rstd).float()
mean, rstd = mean.float(), rstd.float()
del xd
p = torch.empty_like(dp)
xm = torch.empty_like(dp)
idx = torch.empty(dp.shape, dtype=torch.uint8, device=dev)
dx = torch.empty_like(x)
s1 = torch.zeros(Cc, dtype=torch.float64, device=dev)
s2 = torch.zeros(Cc, dtype=torch.float64, device=dev)
n1 = torch.zeros(Cc, dtype=torch.float64, device=dev)
n2 = torch.zeros(Cc, dtype=torch.float64, device=dev)
dgam, dbet = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
pg = train.BnParamGrads()
pg.dgamma, pg.dbeta, pg.n, pg.accumulate = dgam.data_ptr(), dbet.data_ptr(), Cc, 0
vx, vp, vdp, vdx = view4(x), view4(p), view4(dp), view4(dx)
T = x.numel() * 2


def fwd():
    check(lib().pmoe_bn_relu_maxpool_fwd(C.byref(vx), scale.data_ptr(), shift.data_ptr(), C.byref(vp), idx.data_ptr(), xm.data_ptr(),
                                         stream_ptr()), "fwd")


def reduce():
    check(lib().pmoe_bn_relu_maxpool_bwd_reduce(C.byref(vdp), xm.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                                rstd.data_ptr(), s1.data_ptr(), s2.data_ptr(), None, stream_ptr()), "reduce")


def apply(chain):
    check(lib().pmoe_bn_relu_maxpool_bwd_apply(
        C.byref(vdp), idx.data_ptr(), C.byref(vx), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
        s1.data_ptr(), s2.data_ptr(), 1.0 / (B * H * W), C.byref(vdx), n1.data_ptr() if chain else None, n2.data_ptr() if chain else None,
        C.byref(pg), stream_ptr()), "apply")


cases = [("bn_relu_maxpool_fwd", fwd, T * (1 + 0.25 + 0.25 + 0.125)), ("bn_relu_maxpool_bwd_reduce", reduce, T * 0.5),
         ("bn_relu_maxpool_bwd_apply", lambda: apply(False), T * (0.25 + 0.125 + 1 + 1)),
         ("bn_relu_maxpool_bwd_apply<sums>", lambda: apply(True), T * (0.25 + 0.125 + 1 + 1))]
peak = 6551.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
res = {"B": B, "geometry": "%dx%dx%d bf16 NHWC" % (H, W, Cc), "hbm_peak_gbs": peak, "kernels": []}
for name, fn, nbytes in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rec = {"kernel": name, "ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
    res["kernels"].append(rec)
    print("%-34s %7.3f ms  %6.2f GB  %6.0f GB/s  %.2f of peak" % (name, ms, nbytes / 1e9, rec["gbs"], rec["frac_of_hbm_peak"]))
if out:
    json.dump(res, open(out, "w"), indent=1)
