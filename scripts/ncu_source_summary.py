"""Summarise an `ncu --page source --csv` export: per-opcode executed instructions / shared-memory wavefronts /
stall samples and the hottest SASS lines. Usage: python scripts/ncu_source_summary.py file.csv [kernel substring]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
# the export holds one block per kernel launch: "Kernel Name" row, header row, data rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
b = [x for x in blocks if want in x["name"]][0]
hdr, data = b["hdr"], b["data"]
idx = {h: i for i, h in enumerate(hdr)}
print(b["name"], len(data), "SASS lines")
agg = defaultdict(lambda: [0, 0, 0])
tot = 0


def num(r, k):
    try:
        return int(float(r[idx[k]]))
    except Exception:
        return 0


for r in data:
    src = r[idx["Source"]].strip().split()
    if not src:
        continue
    op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
    op = op.split(".")[0]
    agg[op][0] += num(r, "Instructions Executed")
    agg[op][1] += num(r, "L1 Wavefronts Shared")
    agg[op][2] += num(r, "# Samples")
    tot += num(r, "# Samples")
for op, (ex, wf, smp) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:22]:
    print("%-12s exec %10d  smem_wavefronts %10d  samples %6d (%.1f%%)" % (op, ex, wf, smp, 100.0 * smp / max(tot, 1)))
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:28]:
    s = sorted(((h, num(r, h)) for h in stalls), key=lambda kv: -kv[1])[:2]
    print(str(num(r, "# Samples")).rjust(6), r[idx["Source"]].strip()[:72].ljust(72), s)
