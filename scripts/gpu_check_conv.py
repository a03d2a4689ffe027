"""Round-1 GPU bring-up of the tcgen05 conv kernel + the UMMA descriptor-view probe.
Run on the GPU box: python scripts/gpu_check_conv.py [--perf]. Writes gpurun_out/conv_check.json."""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pmoe_b200 import _lib, ops

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
report = {"probe": [], "conv": [], "perf": []}


def probe():
    g = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randn(384, 64, generator=g).to(torch.bfloat16).to(dev)
    B = torch.randn(64, 64, generator=g).to(torch.bfloat16).to(dev)
    Af, Bf = A.float(), B.float()
    variants = [(0, 8, 0), (8, 8, 0), (16, 8, 0), (1, 8, 0), (1, 8, 1), (2, 8, 1), (3, 8, 1), (7, 8, 1),
                (0, 16, 0), (1, 16, 1), (1, 16, 0), (3, 16, 1), (0, 10, 0), (0, 10, 1), (1, 10, 1), (1, 10, 0),
                (11, 10, 1), (11, 10, 0), (22, 10, 1), (5, 18, 1), (5, 18, 0), (0, 9, 1), (0, 9, 0)]
    for (start, grp, bo) in variants:
        D = torch.zeros(128, 64, device=dev)
        rc = _lib.lib().pmoe_dbg_umma_view(A.data_ptr(), 384, B.data_ptr(), D.data_ptr(), start, grp, bo, _lib.stream_ptr())
        _lib.check(rc, "dbg")
        torch.cuda.synchronize()
        rows = torch.tensor([start + (m // 8) * grp + (m % 8) for m in range(128)], device=dev)
        ref = Af[rows] @ Bf.t()
        err = (D - ref).abs().max().item()
        report["probe"].append({"start": start, "group_rows": grp, "bo_mode": bo, "max_err": err, "ok": err < 0.05})
        print("probe start=%2d group=%2d bo=%d  max_err=%.4g" % (start, grp, bo, err), flush=True)


def to_nhwc_pad(x, cpad):
    n, c, h, w = x.shape
    out = torch.zeros(n, h, w, cpad, dtype=torch.bfloat16, device=x.device)
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def run_conv(name, B, H, W, cins, cout, k=3, act="relu", affine=True, residual=False, stats=False, pool=False,
             perf=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    xs = [torch.randn(B, c, H, W, generator=g).to(dev) for c in cins]
    cin = sum(cins)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    scale = (torch.rand(cout, generator=g) + 0.5).to(dev) if affine else None
    shift = (torch.randn(cout, generator=g) * 0.1).to(dev) if affine else None
    cpads = [ops.pad_ch(c) for c in cins]
    ck = ops.choose_ck(cpads)
    cop = ops.cout_padded(cout)
    pad = k // 2
    taps = [(r, s) for r in range(k) for s in range(k)]
    wp = ops.pack_conv_weight(w, cins, cpads, taps, cop)
    segs = ops.conv_segments([(r - pad, s - pad) for (r, s) in taps], cpads, ck)
    srcs = [to_nhwc_pad(x, cp) for x, cp in zip(xs, cpads)]
    cstore = ops.pad_ch(cout)
    out = torch.full((B, H, W, cstore), 7.0, dtype=torch.bfloat16, device=dev)
    res = None
    resf = None
    if residual:
        resf = torch.randn(B, cout, H, W, generator=g).to(dev)
        res = to_nhwc_pad(resf, cstore)
    ssum = torch.zeros(cop, device=dev, dtype=torch.float64) if stats else None
    ssq = torch.zeros(cop, device=dev, dtype=torch.float64) if stats else None
    psum = torch.zeros(B, cop, device=dev) if pool else None
    sc = ops.pad_vec(scale, cop, 1.0) if affine else None
    sh = ops.pad_vec(shift, cop, 0.0) if affine else None
    ops.conv_tc(srcs, wp, segs, ck, out, sc, sh, act, res, ssum, ssq, psum)
    torch.cuda.synchronize()
    # reference on bf16-rounded operands, fp32 math
    xr = torch.cat([x.to(torch.bfloat16).float() for x in xs], 1)
    wr = w.to(torch.bfloat16).float()
    raw = F.conv2d(xr, wr, padding=pad)
    y = raw
    if affine:
        y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if residual:
        y = y + resf.to(torch.bfloat16).float()
    if act == "relu":
        y = torch.relu(y)
    elif act == "elu":
        y = F.elu(y)
    got = out[..., :cout].permute(0, 3, 1, 2).float()
    err = (got - y).abs().max().item()
    rel = ((got - y).norm() / (y.norm() + 1e-12)).item()
    padz = out[..., cout:].float().abs().max().item() if cstore > cout else 0.0
    rec = {"name": name, "B": B, "H": H, "W": W, "cins": cins, "cout": cout, "k": k, "ck": ck, "cout_pad": cop,
           "max_err": err, "rel_err": rel, "pad_max": padz}
    if stats:
        rs = raw.sum(dim=(0, 2, 3))
        rq = (raw * raw).sum(dim=(0, 2, 3))
        rec["stat_sum_rel"] = ((ssum[:cout].float() - rs).norm() / (rs.norm() + 1e-12)).item()
        rec["stat_sq_rel"] = ((ssq[:cout].float() - rq).norm() / (rq.norm() + 1e-12)).item()
    if pool:
        rp = got.sum(dim=(2, 3))
        rec["pool_rel"] = ((psum[:, :cout] - rp).norm() / (rp.norm() + 1e-12)).item()
    rec["ok"] = bool(rel < 1e-2 and padz == 0.0 and rec.get("stat_sum_rel", 0) < 1e-3 and rec.get("stat_sq_rel", 0) < 1e-3
                     and rec.get("pool_rel", 0) < 1e-3)
    if perf:
        for _ in range(3):
            ops.conv_tc(srcs, wp, segs, ck, out, sc, sh, act, res, None, None, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.conv_tc(srcs, wp, segs, ck, out, sc, sh, act, res, None, None, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        flops = 2.0 * B * H * W * cout * cin * k * k
        rec["ms"] = ms
        rec["tflops"] = flops / ms / 1e9
    print(json.dumps(rec), flush=True)
    report["conv"].append(rec)
    return rec


def main():
    perf = "--perf" in sys.argv
    _lib.check(_lib.lib().pmoe_device_check(), "device_check")
    if "--noprobe" not in sys.argv:
        probe()
    run_conv("c64_32", 2, 32, 32, [64], 64)
    run_conv("c64_noaff", 1, 16, 16, [64], 64, act=None, affine=False)
    run_conv("c128_16", 2, 16, 16, [128], 128)
    run_conv("c256_16", 2, 16, 16, [256], 256)
    run_conv("c512_14", 2, 14, 14, [512], 512)
    run_conv("c3_32", 2, 32, 32, [3], 64)
    run_conv("cat64_64", 2, 32, 32, [64, 64], 64)
    run_conv("c92_24", 2, 24, 24, [92], 64)
    run_conv("c1x1_23", 2, 32, 32, [64], 23, k=1, act=None)
    run_conv("c64_3", 2, 32, 32, [64], 3)
    run_conv("linear", 1, 1, 300, [1536], 512, k=1, act="elu")
    run_conv("stats_pool", 3, 28, 28, [64], 128, stats=True, pool=True)
    run_conv("stats_partial", 2, 14, 14, [128], 64, stats=True, pool=True)
    run_conv("residual", 2, 28, 28, [64], 64, residual=True)
    run_conv("c12_56", 2, 56, 56, [12], 64)
    run_conv("c1024_28", 1, 28, 28, [512, 512], 512)
    if perf:
        run_conv("perf_64_224", 32, 224, 224, [64], 64, perf=True)
        run_conv("perf_128_112", 32, 112, 112, [128], 128, perf=True)
        run_conv("perf_256_56", 32, 56, 56, [256], 256, perf=True)
        run_conv("perf_512_28", 32, 28, 28, [512], 512, perf=True)
        run_conv("perf_cat_224", 16, 224, 224, [64, 64], 64, perf=True)
        run_conv("perf_1024_28", 32, 28, 28, [512, 512], 512, perf=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(report, open("gpurun_out/conv_check.json", "w"), indent=1)
    bad = [r["name"] for r in report["conv"] if not r["ok"]]
    print("FAILED:" if bad else "ALL CONV OK", bad)


if __name__ == "__main__":
    main()
