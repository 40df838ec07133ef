#!/bin/bash
# Round 2, twelfth GPU call: same-box A/B of the session-start library, the current one with and without look-ahead draws; K3 v4; ncu for per-line counts.
tag=${1:-r02l}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-FourRooms-v0 MiniGrid-LavaCrossingS9N1-v0; do
  echo "$env base: $(MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_base.so $B --env $env 2>/dev/null | line)"
  echo "$env cur la=0: $(MINIGRID_B200_LOOKAHEAD=0 $B --env $env 2>/dev/null | line)"
  echo "$env cur la=1: $(MINIGRID_B200_LOOKAHEAD=1 $B --env $env 2>$out/la1_$env.err | line)"; tail -3 $out/la1_$env.err
  echo "$env base: $(MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_base.so $B --env $env 2>/dev/null | line)"
done
echo "--- K3"
echo base; MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_base.so timeout 120 python scripts/k3_time.py 2>&1 | tail -3
echo cur; timeout 120 python scripts/k3_time.py 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "wrapper_classes or both_hbm_layouts or (lockstep_vs_oracle and (Empty-5x5 or DistShift2 or FourRooms or MultiRoom-N2 or Dynamic-Obstacles-5x5))" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_$tag.log
echo "--- ncu"
for la in 0 1; do
  MINIGRID_B200_LOOKAHEAD=$la timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $out/${tag}_prof_doorkey_la$la \
    python bench.py --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full la=$la rc=$?"
done
