#!/bin/bash
# Round 2, eighth GPU call: ObstructedMaze (Box.contains), RNG prefetch, host-path trace.
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
echo "--- e2e probe"
timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12
echo "--- timeline"
MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-DoorKey-8x8-v0 262144 2>&1 | tail -22 | head -12
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
for env in MiniGrid-ObstructedMaze-1Dlhb-v0 MiniGrid-ObstructedMaze-Full-v1; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 10 > $out/bench_${tag}_$env.json 2>/dev/null
  echo "$env: $(python -c "import json;d=json.load(open('$out/bench_${tag}_$env.json'));print(d['value'], d['roofline']['frac'], d['run']['autoreset_fraction_per_step'])")"
done
