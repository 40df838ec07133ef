#!/bin/bash
# Round 2, fifth GPU call: idle-lane template fill + dense threshold 8, K3 with batched loads, device wrappers.
tag=${1:-r02e}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
echo "--- timelines"
MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-DoorKey-8x8-v0 262144 2>&1 | tail -22
TL_MODE=2 MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-FourRooms-v0 262144 2>&1 | tail -44
TL_MODE=2 MINIGRID_B200_PDL=0 MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-FourRooms-v0 262144 2>&1 | tail -22 | head -12
TL_MODE=0 MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-LavaCrossingS9N1-v0 262144 2>&1 | tail -22
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
echo "--- sweeps"
for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-LavaCrossingS9N1-v0 MiniGrid-FourRooms-v0 MiniGrid-MultiRoom-N6-v0; do
  echo "$env: $(timeout 100 python scripts/size_sweep.py $env 262144 2>&1 | tail -1)"
done
for env in MiniGrid-MultiRoom-N6-v0 MiniGrid-GoToDoor-8x8-v0 MiniGrid-Playground-v0 MiniGrid-DoorKey-16x16-v0; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 10 > $out/bench_${tag}_$env.json 2>/dev/null
  echo "$env: $(python -c "import json;d=json.load(open('$out/bench_${tag}_$env.json'));print(d['value'], d['roofline']['frac'], d['run']['autoreset_fraction_per_step'])")"
done
