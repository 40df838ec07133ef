#!/bin/bash
# Round 2, thirteenth GPU call: look-ahead draws removed, reward wrappers out of line; same-box A/B against the session-start library;
# full parity suite (with the many-tiles-per-warp test); bench with the e2e leg fixed (torch CPU threads off, 3 repetitions).
tag=${1:-r02m}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-FourRooms-v0 MiniGrid-LavaCrossingS9N1-v0; do
  echo "$env base: $(MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_base.so $B --env $env 2>/dev/null | line)"
  echo "$env cur : $($B --env $env 2>/dev/null | line)"
  echo "$env base: $(MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_base.so $B --env $env 2>/dev/null | line)"
  echo "$env cur : $($B --env $env 2>/dev/null | line)"
done
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
echo "--- bench"
timeout 600 python bench.py --no-cpu-baseline > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -3 $out/bench_$tag.err
python - <<PY
import json
d = json.load(open("$out/bench_$tag.json"))
print(d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"))
PY
