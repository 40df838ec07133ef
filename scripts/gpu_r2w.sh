#!/bin/bash
# Round 2: one-buffer tiled kernel, next tile pulled at the top of the iteration and its block / records / actions L2-prefetched: A/B and parity.
tag=${1:-r02w}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
for cfg in "MiniGrid-DoorKey-8x8-v0 262144" "MiniGrid-LavaCrossingS9N1-v0 262144" "MiniGrid-Empty-8x8-v0 65536" "MiniGrid-Empty-8x8-v0 262144" "MiniGrid-Fetch-8x8-N3-v0 262144"; do
  set -- $cfg
  echo "$1 x $2 prefetch off: $(MINIGRID_B200_WINPREF=0 $B --env $1 --envs-per-gpu $2 2>/dev/null | line)"
  echo "$1 x $2 prefetch on : $($B --env $1 --envs-per-gpu $2 2>/dev/null | line)"
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "both_hbm_layouts or many_tiles or full_size or autoreset_on_full or packed_host or reward_wrappers_lockstep or (lockstep_vs_oracle and (DoorKey-8x8 or LavaCrossingS9N1 or Empty or Dynamic or Fetch or GoToDoor or Memory))" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_$tag.log
