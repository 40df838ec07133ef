#!/usr/bin/env python
"""Summarise an ncu report of k_step here (no GPU needed): key metrics, stall mix, hottest SASS lines."""
import csv
import subprocess
import sys

rep = sys.argv[1]
n_tiles = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
seen = 0
keys = ("Duration", "Elapsed Cycles", "SM Active Cycles", "Executed Ipc", "Issue Slots Busy", "Registers Per", "Theoretical Active Warps",
        "Achieved Active Warps", "Eligible Warps", "Active Warps Per Sch", "No Eligible", "Executed Instructions", "Grid Size",
        "DRAM Throughput", "Dynamic Shared", "Memory Throughput", "L2 Hit", "way bank")
for line in det.splitlines():
    if "k_step" in line and "Context" in line:
        seen += 1
    if seen == 1 and any(k in line for k in keys):
        print(line.rstrip()[:150])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
    if m in hdr:
        i = hdr.index(m)
        print(m, rows[1][i], [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
idx = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    data.append(r)
tot = sum(int(r[idx["# Samples"]]) for r in data)
stall = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[idx[c]]) for r in data) for c in stall}
print("samples", tot, " ".join(f"{c[6:]}={100 * v / tot:.1f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
ie = sum(int(r[idx["Instructions Executed"]]) for r in data)
print("warp-instructions per tile:", ie / n_tiles)
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:12]:
    st = {c[6:]: int(r[idx[c]]) for c in stall if int(r[idx[c]]) > 0}
    print(f"{data.index(r):5d} {int(r[idx['# Samples']]):5d} {int(r[idx['Instructions Executed']]):7d} {r[idx['Source']].strip()[:56]:56s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")
