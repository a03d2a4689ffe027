"""Repeat the reduction kernels on fixed inputs and look for run-to-run differences beyond summation-order noise."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import _lib, nhwc
from pmoe_b200._lib import check, lib, stream_ptr, view4
from pmoe_b200.nhwc import dtype_code

torch.manual_seed(0)
bad = 0
for dtype in (torch.float32, torch.bfloat16):
    for (n, h, w, c) in [(2, 32, 32, 64), (2, 2, 2, 512), (2, 16, 16, 128), (3, 8, 8, 256), (2, 32, 32, 16), (2, 4, 4, 48)]:
        dz = torch.randn(n, h, w, c, device="cuda").to(dtype)
        z = torch.randn(n, h, w, c, device="cuda").to(dtype)
        x = torch.randn(n, h, w, c, device="cuda").to(dtype)
        mean = torch.randn(c, device="cuda") * 0.1
        rstd = torch.rand(c, device="cuda") + 0.5
        ref = None
        worst = 0.0
        for it in range(300):
            s1 = torch.zeros(c, dtype=torch.float64, device="cuda")
            s2 = torch.zeros(c, dtype=torch.float64, device="cuda")
            vdz, vz, vx = view4(dz), view4(z), view4(x)
            check(lib().pmoe_bn_bwd_reduce(C.byref(vdz), C.byref(vz), C.byref(vx), dtype_code(dz), 1, mean.data_ptr(), rstd.data_ptr(),
                                           s1.data_ptr(), s2.data_ptr(), None, None, stream_ptr()), "reduce")
            st1 = torch.zeros(c, dtype=torch.float64, device="cuda")
            st2 = torch.zeros(c, dtype=torch.float64, device="cuda")
            check(lib().pmoe_channel_stats(C.byref(vx), dtype_code(x), st1.data_ptr(), st2.data_ptr(), stream_ptr()), "stats")
            cs = nhwc.channel_sums(x)
            cur = torch.cat([s1, s2, st1, st2, cs.double().reshape(-1)])
            if ref is None:
                ref = cur.clone()
            else:
                worst = max(worst, ((cur - ref).abs() / (ref.abs() + 1e-3)).max().item())
        # reference values in torch
        d = dz.float() * (z.float() > 0)
        e1 = (s1 - d.double().sum((0, 1, 2))).abs().max().item()
        xh = (x.float() - mean) * rstd
        e2 = (s2 - (d * xh).double().sum((0, 1, 2))).abs().max().item()
        flag = worst > 1e-5 or e1 > 1e-2 or e2 > 1e-2
        bad += flag
        print("%-8s %-18s run-to-run %.2e | vs torch s1 %.2e s2 %.2e %s" % (str(dtype).split(".")[1], (n, h, w, c), worst, e1, e2, "BAD" if flag else ""), flush=True)
print("RACE_CHECK", "FAIL" if bad else "OK")
