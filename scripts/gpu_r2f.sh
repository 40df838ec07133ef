#!/bin/bash
# Round 2, sixth GPU call: RoomGrid on the device, early truncation flags, K3 v3.
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
echo "--- timeline"
MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-DoorKey-8x8-v0 262144 2>&1 | tail -22
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
for env in MiniGrid-Unlock-v0 MiniGrid-BlockedUnlockPickup-v0 MiniGrid-KeyCorridorS6R3-v0 MiniGrid-KeyCorridorS3R3-v0; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 10 > $out/bench_${tag}_$env.json 2>/dev/null
  echo "$env: $(python -c "import json;d=json.load(open('$out/bench_${tag}_$env.json'));print(d['value'], d['roofline']['frac'], d['run']['autoreset_fraction_per_step'])")"
done
