#!/bin/bash
# Round 2, fourth GPU call: order list built in the PDL prologue (regenerating tiles in a CTA's first round), K3 v2.
tag=${1:-r02d}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
echo "--- timelines"
for H in 1 0; do
  echo "hot_first=$H"; MINIGRID_B200_HOTFIRST=$H MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-DoorKey-8x8-v0 262144 2>&1 | tail -14 | head -32
done
TL_MODE=2 MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-FourRooms-v0 262144 2>&1 | tail -28
TL_MODE=0 MINIGRID_B200_VERBOSE=1 MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-LavaCrossingS9N1-v0 262144 2>&1 | tail -30
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
MINIGRID_B200_HOTFIRST=0 timeout 300 python bench.py --no-cpu-baseline --no-configs > $out/bench_${tag}_nohot.json 2>/dev/null
python - <<PY
import json
for f in ("$out/bench_$tag.json", "$out/bench_${tag}_nohot.json"):
    d = json.load(open(f))
    print(f, d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
echo "--- sweeps"
for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-LavaCrossingS9N1-v0 MiniGrid-FourRooms-v0; do
  echo "$env: $(timeout 100 python scripts/size_sweep.py $env 262144 2>&1 | tail -1)"
done
