#!/bin/bash
# Round 2, third GPU call: Dynamic-Obstacles, packed fix; where does a desynchronised step lose its time (timeline)?
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_$tag.log
echo "--- host expansion"
timeout 120 python scripts/expand_bench.py
echo "--- timelines"
for H in 1 0; do
  echo "hot_first=$H"; MINIGRID_B200_HOTFIRST=$H MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-DoorKey-8x8-v0 262144 2>&1 | tail -30
done
MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-FourRooms-v0 262144 2>&1 | tail -30
MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-LavaCrossingS9N1-v0 262144 2>&1 | tail -30
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
for env in MiniGrid-Dynamic-Obstacles-8x8-v0 MiniGrid-Dynamic-Obstacles-16x16-v0; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 10 > $out/bench_${tag}_$env.json 2>/dev/null
  echo "$env: $(python -c "import json;d=json.load(open('$out/bench_${tag}_$env.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'])")"
done
