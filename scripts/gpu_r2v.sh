#!/bin/bash
# Round 2: window layout, L2 prefetch of the next tile's view lines (mid-tile, from the records that have arrived): A/B and parity.
tag=${1:-r02v}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
for env in MiniGrid-FourRooms-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-LockedRoom-v0 MiniGrid-ObstructedMaze-Full-v1; do
  echo "$env prefetch off: $(MINIGRID_B200_WINPREF=0 $B --env $env 2>/dev/null | line)"
  echo "$env prefetch on : $($B --env $env 2>/dev/null | line)"
done
echo "DoorKey (tiled, unaffected): $($B 2>/dev/null | line)"
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "both_hbm_layouts or many_tiles or full_size or autoreset_on_full or roomgrid_post or (lockstep_vs_oracle and (FourRooms or MultiRoom or 16x16 or KeyCorridor or ObstructedMaze-Full-v1 or LockedRoom or Playground))" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_$tag.log
