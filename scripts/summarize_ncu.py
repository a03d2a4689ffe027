"""Turn ncu exports in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_punet_b64.json
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_conv_halo64_full.json [kernel substring]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]


def launches(src, dst):
    hdr, rows = None, []
    for r in csv.reader(open(src)):
        if r and r[0] == "ID":
            hdr = r
        elif hdr and r and r[0].isdigit():
            rows.append(r)
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r[i_name]).replace("void ", "")
        v = float(r[i_val].replace(",", ""))
        v = v / 1000.0 if r[i_unit] == "ns" else v
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    out = {"source": src, "note": "ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised launches; compare SHARES",
           "launches": len(rows), "total_us": tot,
           "kernels": [{"kernel": k, "us": v, "share": v / tot, "launches": n} for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])]}
    json.dump(out, open(dst, "w"), indent=1)
    for k in out["kernels"]:
        print("%-56s %9.1f us %5.1f%% n=%d" % (k["kernel"][:56], k["us"], 100 * k["share"], k["launches"]))


def full(src, dst, want=""):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"source": src, "launches": []}
    for r in rows[2:]:
        if want and want not in r[idx["Kernel Name"]]:
            continue
        rec = {"kernel": r[idx["Kernel Name"]]}
        for k in KEYS:
            if k in idx:
                rec[k] = {"value": r[idx[k]], "unit": units[idx[k]]}
        out["launches"].append(rec)
    json.dump(out, open(dst, "w"), indent=1)
    for rec in out["launches"]:
        print(rec["kernel"])
        for k in KEYS:
            if k in rec:
                print("   %-90s %s %s" % (k, rec[k]["value"], rec[k]["unit"]))


def perkernel(src, dst):
    """One record per kernel instantiation of a --set full capture: its LONGEST launch (the representative big tensor),
    with duration, DRAM bytes / throughput, tensor-pipe and issue utilisation, registers, grid."""
    if src.endswith(".csv"):  # `ncu -i x.ncu-rep --page raw --csv > x.csv` made on the GPU box (reports can exceed the 64 MiB pull limit)
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.splitlines()) if r]
    while rows and rows[0][0] != "ID":
        rows.pop(0)
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    best, count = {}, collections.Counter()

    def num(r, k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except Exception:
            return 0.0
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        count[name] += 1
        if name not in best or num(r, "gpu__time_duration.sum") > num(best[name], "gpu__time_duration.sum"):
            best[name] = r
    out = {"source": src, "note": "ncu --set full --clock-control none; per kernel instantiation the longest captured launch",
           "kernels": []}
    for name, r in sorted(best.items(), key=lambda kv: -num(kv[1], "gpu__time_duration.sum")):
        rec = {"kernel": name, "captured_launches": count[name]}
        for k in KEYS:
            if k in idx:
                rec[k] = {"value": r[idx[k]], "unit": units[idx[k]]}
        out["kernels"].append(rec)
        print("%-60s %9s %s  dram %5s%%  tensor %5s%%  sm %5s%%" % (
            name[:60], r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]],
            r[idx["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]] if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in idx else "?",
            r[idx[KEYS[4]]] if KEYS[4] in idx else "?",
            r[idx["sm__throughput.avg.pct_of_peak_sustained_elapsed"]] if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in idx else "?"))
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "perkernel":
        perkernel(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
