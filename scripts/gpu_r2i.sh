#!/bin/bash
# Round 2, ninth GPU call (session 2): confirm the restored tree (process-wide host pool, alternating cold order) on the GPU.
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
