#!/bin/bash
# Round 2, tenth GPU call: window mode with the two-deep cp.async pipeline (parity, plan sweep, timeline, ncu), e2e trace in bench context.
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
B="python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 600 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
echo "--- FourRooms plan sweep (warps,vis,nbuf)"
for cfg in auto 14,2,1 16,2,1 17,2,1 12,2,1 19,1,1 20,1,1 16,1,1; do
  if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
  echo "$cfg: $(timeout 120 $B --env MiniGrid-FourRooms-v0 2>/dev/null | line)"
done
unset MINIGRID_B200_CFG
for env in MiniGrid-MultiRoom-N6-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-KeyCorridorS6R3-v0 MiniGrid-ObstructedMaze-Full-v1 MiniGrid-LockedRoom-v0; do
  echo "$env: $(timeout 120 $B --env $env 2>/dev/null | line)"
done
echo "--- timeline FourRooms"
TL_MODE=2 MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 120 python scripts/timeline2.py MiniGrid-FourRooms-v0 262144 2>&1 | tail -14
echo "--- e2e in bench context"
MINIGRID_B200_HOST_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 200 --warmup 10 > $out/bench_${tag}_e2e.json 2> $out/bench_${tag}_e2e.err; tail -6 $out/bench_${tag}_e2e.err; python -c "import json;d=json.load(open('$out/bench_${tag}_e2e.json'));print(d['e2e'])"
MINIGRID_B200_HOST_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --no-configs --steps 200 --warmup 10 --rotate 2 2> $out/bench_${tag}_e2e_r2.err | python -c "import json,sys;d=json.loads(sys.stdin.read());print('rotate 2:', d['e2e'])"; tail -3 $out/bench_${tag}_e2e_r2.err
timeout 300 python scripts/e2e_probe.py 2>&1 | sed -n 1,5p
echo "--- ncu"
for w in fourrooms:MiniGrid-FourRooms-v0 doorkey:MiniGrid-DoorKey-8x8-v0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o $out/${tag}_prof_${w%%:*} \
    python bench.py --env ${w##*:} --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full ${w%%:*} rc=$?"
done
ls -la $out/*.ncu-rep
