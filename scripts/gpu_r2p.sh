#!/bin/bash
# Round 2, sixteenth GPU call: TILED1 with a 28-warp / 72-register bound: plan sweep over the tiled kinds; e2e calibration in blocks.
tag=${1:-r02p}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
sweep() { # env, n, cfgs...
  env=$1; n=$2; shift 2
  for cfg in "$@"; do
    if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
    echo "$env x $n $cfg: $($B --env $env --envs-per-gpu $n 2>/dev/null | line)"
  done
  unset MINIGRID_B200_CFG
}
sweep MiniGrid-DoorKey-8x8-v0 262144 auto 14,2,2 20,2,1 22,2,1 24,2,1 26,2,1 28,2,1 28,1,1
sweep MiniGrid-Empty-8x8-v0 65536 auto 14,0,2 14,0,1 20,0,1 28,0,1
sweep MiniGrid-Empty-8x8-v0 262144 auto 14,0,2 24,0,1 28,0,1
sweep MiniGrid-LavaCrossingS9N1-v0 262144 auto 18,2,1 20,2,1
sweep MiniGrid-GoToDoor-8x8-v0 262144 auto 28,0,1 14,0,2
sweep MiniGrid-DoorKey-5x5-v0 262144 auto 28,2,1 14,2,2
sweep MiniGrid-LavaGapS7-v0 262144 auto 28,2,1 14,2,2
sweep MiniGrid-Dynamic-Obstacles-8x8-v0 262144 auto 28,0,1 14,0,2
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 -k "many_tiles or full_size or packed_host or (lockstep_vs_oracle and (DoorKey-8x8 or LavaCrossingS9N1 or Empty-8x8))" > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_$tag.log
echo "--- e2e"
E="timeout 200 python bench.py --no-cpu-baseline --no-configs --steps 200 --warmup 10"
e2e() { python -c "import json,sys;d=json.loads(sys.stdin.read());print(['%.3g'%v for v in d['e2e']['repetitions']], 'full', ['%.3g'%v for v in d['e2e']['full_format']['repetitions']])" 2>&1 | tail -1; }
echo "R=4 auto : $(MINIGRID_B200_HOST_TRACE=1 $E 2>$out/e2e_$tag.err | e2e)"; grep "host expansion" $out/e2e_$tag.err | head -4; tail -1 $out/e2e_$tag.err
echo "R=1 auto : $(MINIGRID_B200_HOST_TRACE=1 $E --rotate 1 2>$out/e2e1_$tag.err | e2e)"; grep "host expansion" $out/e2e1_$tag.err | head -2; tail -1 $out/e2e1_$tag.err
