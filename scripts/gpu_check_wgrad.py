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


def run_general(name, srcs, segs, ck, dy, cop, perf=False, flops=0.0):
    """TC vs SIMT on an arbitrary descriptor (sources may be strided views)."""
    ktot = ck * sum(sg[4] for sg in segs)
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
    rec = {"name": name, "supported": True, "rel_vs_simt": err, "max_abs": (dw_tc - dw_si).abs().max().item(), "rel_vs_torch": None}
    if perf:
        for fn_name, call, n in (("tc", lambda: _lib.lib().pmoe_conv_wgrad_tc(C.byref(d1), dw_tc.data_ptr(), _lib.stream_ptr()), 5),
                                 ("simt", lambda: _lib.lib().pmoe_conv_wgrad_simt(C.byref(d2), _lib.BF16, dw_si.data_ptr(), _lib.stream_ptr()), 2)):
            call()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                call()
            e1.record()
            torch.cuda.synchronize()
            rec[fn_name + "_ms"] = e0.elapsed_time(e1) / n
            rec[fn_name + "_tflops"] = flops / rec[fn_name + "_ms"] / 1e9
    print("%-28s rel vs simt %.3e  max abs %.3e  %s" % (name, err, rec["max_abs"],
          " ".join("%s=%.3g" % (k, v) for k, v in rec.items() if k.endswith(("_ms", "_tflops")))), flush=True)
    report.append(rec)


def rnd(shape, g, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def general_cases(perf):
    g = torch.Generator().manual_seed(5)
    # 3x3 on a 14x14 image (too small for the halo kernel)
    B = 64 if perf else 3
    x = rnd((B, 14, 14, 512), g)
    dy = rnd((B, 14, 14, 512), g, 0.1)
    segs = ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], [512], 64)
    run_general("512->512 14x14 3x3", [x], segs, 64, dy, 512, perf, 2.0 * B * 196 * 512 * 512 * 9)
    # 1x1
    x = rnd((2, 28, 28, 64), g)
    dy = rnd((2, 28, 28, 128), g, 0.1)
    run_general("64->128 28x28 1x1", [x], ops.conv_segments([(0, 0)], [64], 64), 64, dy, 128)
    # stride-2 3x3 over the four parity views (train.stride2_sources)
    B = 64 if perf else 2
    x = rnd((B, 56, 56, 64), g)
    dy = rnd((B, 28, 28, 128), g, 0.1)
    views, vidx, segs = [], {}, []
    par = {0: (1, -1), 1: (0, 0), 2: (1, 0)}
    for r in range(3):
        for s_ in range(3):
            (pp, dh), (qq, dw) = par[r], par[s_]
            if (pp, qq) not in vidx:
                vidx[(pp, qq)] = len(views)
                views.append(x[:, pp::2, qq::2, :])
            segs.append((vidx[(pp, qq)], dh, dw, 0, 1))
    run_general("64->128 s2 3x3 56->28", views, segs, 64, dy, 128, perf, 2.0 * B * 784 * 64 * 128 * 9)
    # 16-channel source (the ResNet stem's first conv: 12 -> 64)
    B = 32 if perf else 2
    hw = 224 if perf else 32
    x = torch.zeros(B, hw, hw, 16, dtype=torch.bfloat16, device=dev)
    x[..., :12] = rnd((B, hw, hw, 12), g)
    dy = rnd((B, hw, hw, 64), g, 0.1)
    segs = ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], [16], 16)
    run_general("12->64 3x3 ck16", [x], segs, 16, dy, 64, perf, 2.0 * B * hw * hw * 12 * 64 * 9)
    # 32-channel chunks (96-channel source)
    x = rnd((2, 32, 32, 96), g)
    dy = rnd((2, 32, 32, 64), g, 0.1)
    run_general("96->64 3x3 ck32", [x], ops.conv_segments([(r - 1, s - 1) for (r, s) in ops.TAPS3], [96], 32), 32, dy, 64)
    # linear layer over a virtual concat: rows = batch
    Bv = 8192 if perf else 300
    srcs = [rnd((1, 1, Bv, 512), g) for _ in range(3)]
    dy = rnd((1, 1, Bv, 512), g, 0.1)
    run_general("linear 1536->512", srcs, ops.conv_segments([(0, 0)], [512, 512, 512], 64), 64, dy, 512, perf, 2.0 * Bv * 1536 * 512)
    # ConvTranspose2d k2s2 weight gradient: dy read through a pixel-shuffle view
    x = rnd((2, 14, 14, 128), g)
    dyf = rnd((2, 28, 28, 64), g, 0.1)
    run_general("convT 128->64 (dy view)", [x], ops.conv_segments([(0, 0)], [128], 64), 64, dyf[:, 1::2, 0::2, :], 64)


if __name__ == "__main__":
    perf = "--perf" in sys.argv
    run("64->64 32x32 B2", 2, 32, 32, [64], 64)
    run("64->64 48x40 B1 (partial)", 1, 48, 40, [64], 64)
    run("128->128 32x32 B2", 2, 32, 32, [128], 128)
    run("64+64->64 32x32 B2 (concat)", 2, 32, 32, [64, 64], 64)
    run("256->256 32x32 B1", 1, 32, 32, [256], 256)
    run("128->23 32x32 B1", 1, 32, 32, [128], 23)
    run("512->512 16x16 B2", 2, 16, 16, [512], 512)
    general_cases(perf)
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
