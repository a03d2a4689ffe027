"""GPU check of the tensor-core weight-gradient kernel against the CUDA-core one (same packed layout) and against
torch's fp32 conv weight gradient on bf16-rounded operands. Writes gpurun_out/wgrad_check.json.
Run on the GPU box: python scripts/gpu_check_wgrad.py [--perf]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import _lib, ops, profiler

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
report = []


def desc_for(srcs, segs, ck, dy, dwp):
    return ops._fill_desc(srcs, dwp, segs, ck, dy, None, None, None, None, None, None, None, None, 0, torch.bfloat16)


def run(name, B, H, W, cins, cout, perf=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    cpads = [ops.pad_ch(c) for c in cins]
    ck = ops.choose_ck(cpads)
    cop, cstore = ops.cout_padded(cout), ops.pad_ch(cout)
    srcs = []
    for c, cp in zip(cins, cpads):
        t = torch.zeros(B, H, W, cp, dtype=torch.bfloat16, device=dev)
        t[..., :c] = torch.randn(B, H, W, c, generator=g).to(torch.bfloat16).to(dev)
        srcs.append(t)
    dy = torch.zeros(B, H, W, cstore, dtype=torch.bfloat16, device=dev)
    dy[..., :cout] = (torch.randn(B, H, W, cout, generator=g) * 0.1).to(torch.bfloat16).to(dev)
    segs = ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], cpads, ck)
    ktot = 9 * sum(cpads)
    dw_tc = torch.zeros(cop, ktot, dtype=torch.float32, device=dev)
    dw_si = torch.zeros(cop, ktot, dtype=torch.float32, device=dev)
    d1 = desc_for(srcs, segs, ck, dy, dw_tc)
    rc = _lib.lib().pmoe_conv_wgrad_tc(C.byref(d1), dw_tc.data_ptr(), _lib.stream_ptr())
    if rc == -2:
        print("%-28s unsupported by the tensor-core kernel" % name, flush=True)
        report.append({"name": name, "supported": False})
        return
    _lib.check(rc, "wgrad_tc")
    d2 = desc_for(srcs, segs, ck, dy, dw_si)
    _lib.check(_lib.lib().pmoe_conv_wgrad_simt(C.byref(d2), _lib.BF16, dw_si.data_ptr(), _lib.stream_ptr()), "wgrad_simt")
    torch.cuda.synchronize()
    err = ((dw_tc - dw_si).norm() / dw_si.norm()).item()
    mx = (dw_tc - dw_si).abs().max().item()
    # independent check on a small case: torch autograd weight gradient in fp32
    terr = None
    if B * H * W <= 64 * 64 * 8:
        x = torch.cat([s[..., :c].float() for s, c in zip(srcs, cins)], 3).permute(0, 3, 1, 2).contiguous()
        w = torch.zeros(cout, sum(cins), 3, 3, device=dev, requires_grad=True)
        y = torch.nn.functional.conv2d(x, w, None, 1, 1)
        y.backward(dy[..., :cout].float().permute(0, 3, 1, 2).contiguous())
        gw = w.grad  # (cout, cin, 3, 3)
        got = torch.zeros_like(gw)
        off = 0
        for t, (r, s) in enumerate(ops.TAPS3):
            ci = 0
            for c, cp in zip(cins, cpads):
                got[:, ci:ci + c, r, s] = dw_tc[:cout, off:off + c]
                ci += c
                off += cp
        terr = ((got - gw).norm() / gw.norm()).item()
    rec = {"name": name, "supported": True, "rel_vs_simt": err, "max_abs": mx, "rel_vs_torch": terr}
    if perf:
        flops = 2.0 * B * H * W * cout * 9 * sum(cins)
        for fn_name, call in (("tc", lambda: _lib.lib().pmoe_conv_wgrad_tc(C.byref(d1), dw_tc.data_ptr(), _lib.stream_ptr())),
                              ("simt", lambda: _lib.lib().pmoe_conv_wgrad_simt(C.byref(d2), _lib.BF16, dw_si.data_ptr(), _lib.stream_ptr()))):
            for _ in range(2):
                call()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5 if fn_name == "tc" else 2
            e0.record()
            for _ in range(n):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            rec[fn_name + "_ms"] = ms
            rec[fn_name + "_tflops"] = flops / ms / 1e9
    print("%-28s rel vs simt %.3e  max abs %.3e  vs torch %s  %s" % (
        name, err, mx, "%.3e" % terr if terr is not None else "-",
        " ".join("%s=%.3g" % (k, v) for k, v in rec.items() if k.endswith(("_ms", "_tflops")))), flush=True)
    report.append(rec)


if __name__ == "__main__":
    perf = "--perf" in sys.argv
    run("64->64 32x32 B2", 2, 32, 32, [64], 64)
    run("64->64 48x40 B1 (partial)", 1, 48, 40, [64], 64)
    run("128->128 32x32 B2", 2, 32, 32, [128], 128)
    run("64+64->64 32x32 B2 (concat)", 2, 32, 32, [64, 64], 64)
    run("256->256 32x32 B1", 1, 32, 32, [256], 256)
    run("128->23 32x32 B1", 1, 32, 32, [128], 23)
    run("512->512 16x16 B2", 2, 16, 16, [512], 512)
    if perf:
        run("64->64 224x224 B32", 32, 224, 224, [64], 64, perf=True)
        run("128->128 112x112 B32", 32, 112, 112, [128], 128, perf=True)
        run("128+128->128 112 B32", 32, 112, 112, [128, 128], 128, perf=True)
        run("256->256 56x56 B32", 32, 56, 56, [256], 256, perf=True)
        run("512->512 28x28 B32", 32, 28, 28, [512], 512, perf=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/wgrad_check.json", "w"), indent=1)
    bad = [r for r in report if r.get("supported") and r["rel_vs_simt"] > 2e-3]
    print("WGRAD_TC %s" % ("FAIL %r" % [b["name"] for b in bad] if bad else "OK"))
    sys.exit(1 if bad else 0)
