#!/bin/bash
# Round 2, last call: smoke + the default bench line of the final source.
tag=${1:-r02x}
out=gpurun_out
mkdir -p $out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -2 $out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("$out/${tag}_bench.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d["full_obs"]["frac"], d["cpu_baseline"]["value"])
PY
