#!/usr/bin/env python
"""Dynamic warp-instructions per source line: joins ncu's per-SASS-instruction executed counts with nvdisasm -g
line info of the same kernel (same cubin, same instruction order).
usage: ncu_lines.py <report.ncu-rep> <mangled-kernel-prefix> [n_tiles]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, mangled = sys.argv[1], sys.argv[2]
n_tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.environ.get("NCU_LINES_LIB", os.path.join(root, "minigrid_b200", "libminigrid_b200.so"))], cwd=tmp, capture_output=True)
sass = []
for cub in sorted(f for f in os.listdir(tmp) if f.startswith("mg_step")):  # K1 is instantiated in several translation units
    out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    if ".text." + mangled in out:
        sass = out.split("\n")
        break
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + mangled))
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith(".text.") or sass[i].startswith(".section"))
lines = []
cur = None
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        lines.append((cur, re.sub(r"\s+", " ", m.group(1)).strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
idx = {n: i for i, n in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    data.append(r)
print("sass instrs: nvdisasm", len(lines), "ncu", len(data))
import difflib
a_ops = [t.split()[0] if not t.startswith("@") else t.split()[1] for _, t in lines]
b_txt = [re.sub(r"\s+", " ", r[idx["Source"]]).strip() for r in data]
b_ops = [t.split()[0] if not t.startswith("@") else t.split()[1] for t in b_txt]
sm = difflib.SequenceMatcher(a=a_ops, b=b_ops, autojunk=False)
agg = collections.Counter()
samp = collections.Counter()
matched = 0
for blk in sm.get_matching_blocks():
    for k in range(blk.size):
        ln = lines[blk.a + k][0]
        r = data[blk.b + k]
        agg[ln] += int(r[idx["Instructions Executed"]])
        samp[ln] += int(r[idx["# Samples"]])
        matched += 1
print("matched", matched)
files = {}
tot = sum(agg.values())
print("total per tile", tot / n_tiles)
for (f, l), c in agg.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 45):
    if f not in files:
        pth = os.path.join(root, "minigrid_b200", "csrc", f)
        files[f] = open(pth).read().split("\n") if os.path.exists(pth) else None
    text = files[f][l - 1].strip()[:95] if files[f] and 0 < l <= len(files[f]) else ""  # (the source may have moved on since the capture)
    print(f"{c / n_tiles:7.1f} {samp[(f, l)]:5d}  {f}:{l}  {text}")
