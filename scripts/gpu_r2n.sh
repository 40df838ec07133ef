#!/bin/bash
# Round 2, fourteenth GPU call: window layout with lines as wide as the grid (A/B against the session-start library, both layouts
# forced on the small grids), e2e with streaming stores / fewer batches.
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
BASE=$PWD/minigrid_b200/libminigrid_b200_base.so
for env in MiniGrid-FourRooms-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-MultiRoom-N6-v0 MiniGrid-KeyCorridorS6R3-v0; do
  echo "$env base: $(MINIGRID_B200_LIB=$BASE $B --env $env 2>/dev/null | line)"
  echo "$env cur : $($B --env $env 2>/dev/null | line)"
done
for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-LavaCrossingS9N1-v0 MiniGrid-Empty-8x8-v0; do
  echo "$env tiled  cur : $($B --env $env 2>/dev/null | line)"
  echo "$env window base: $(MINIGRID_B200_LAYOUT=1 MINIGRID_B200_LIB=$BASE $B --env $env 2>/dev/null | line)"
  echo "$env window cur : $(MINIGRID_B200_LAYOUT=1 $B --env $env 2>/dev/null | line)"
  for cfg in 16,2,1 20,1,1 14,2,1; do echo "$env window cur cfg=$cfg: $(MINIGRID_B200_CFG=$cfg MINIGRID_B200_LAYOUT=1 $B --env $env 2>/dev/null | line)"; done
done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 -k "both_hbm_layouts or many_tiles or full_size or roomgrid_post or autoreset_on_full or (lockstep_vs_oracle and (FourRooms or MultiRoom or 16x16 or KeyCorridor or ObstructedMaze-Full-v1 or LockedRoom))" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_$tag.log
echo "--- e2e"
E="timeout 200 python bench.py --no-cpu-baseline --no-configs --steps 200 --warmup 10"
e2e() { python -c "import json,sys;d=json.loads(sys.stdin.read());print(['%.3g'%v for v in d['e2e']['repetitions']], 'full', ['%.3g'%v for v in d['e2e']['full_format']['repetitions']])" 2>&1 | tail -1; }
echo "R=4 plain : $($E 2>/dev/null | e2e)"
echo "R=4 stream: $(MINIGRID_B200_EXPAND_STREAM=1 $E 2>/dev/null | e2e)"
echo "R=1 plain : $($E --rotate 1 2>/dev/null | e2e)"
echo "R=1 stream: $(MINIGRID_B200_EXPAND_STREAM=1 $E --rotate 1 2>/dev/null | e2e)"
echo "R=2 stream: $(MINIGRID_B200_EXPAND_STREAM=1 $E --rotate 2 2>/dev/null | e2e)"
