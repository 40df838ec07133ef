#!/bin/bash
# Round 2, eleventh GPU call: look-ahead draws (drawing warp), reward wrappers (NoDeath / bonus) parity, host pool with slices.
tag=${1:-r02k}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
echo "--- look-ahead A/B (0 = inline draws)"
for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-FourRooms-v0 MiniGrid-LavaCrossingS9N1-v0 MiniGrid-Empty-8x8-v0 MiniGrid-GoToDoor-8x8-v0 MiniGrid-MultiRoom-N6-v0 MiniGrid-KeyCorridorS6R3-v0; do
  for la in 0 1; do
    echo "$env lookahead=$la: $(MINIGRID_B200_LOOKAHEAD=$la $B --env $env 2>/dev/null | line)"
  done
done
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
echo "--- e2e probe"
timeout 300 python scripts/e2e_probe.py 2>&1 | sed -n 1,5p
echo "--- timeline"
MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-DoorKey-8x8-v0 262144 2>&1 | tail -14
