#!/bin/bash
# Round 2, seventeenth GPU call: how many warps for the one-buffer tiled kernels (72 registers); the window kernels at 28 warps / 72 registers.
tag=${1:-r02q}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
sweep() { env=$1; n=$2; shift 2
  for cfg in "$@"; do
    if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
    echo "$env x $n $cfg: $($B --env $env --envs-per-gpu $n 2>/dev/null | line)"
  done; unset MINIGRID_B200_CFG; }
sweep MiniGrid-DoorKey-8x8-v0 262144 16,2,1 18,2,1 20,2,1 21,2,1
sweep MiniGrid-Empty-8x8-v0 262144 18,0,1 20,0,1 22,0,1
sweep MiniGrid-Empty-8x8-v0 65536 14,0,1 18,0,1 22,0,1
sweep MiniGrid-GoToDoor-8x8-v0 262144 20,0,1 22,0,1 24,0,1
sweep MiniGrid-DoorKey-5x5-v0 262144 20,2,1 22,2,1 24,2,1
sweep MiniGrid-LavaGapS7-v0 262144 20,2,1 22,2,1 24,2,1
sweep MiniGrid-Dynamic-Obstacles-8x8-v0 262144 20,0,1 22,0,1 24,0,1
sweep MiniGrid-Fetch-8x8-N3-v0 262144 auto 20,0,1 22,0,1 28,0,1
echo "--- window kernels, 28-warp / 72-register build"
export MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_w28.so
sweep MiniGrid-FourRooms-v0 262144 20,2,1 22,2,1 24,2,1 28,2,1 24,1,1 28,1,1
sweep MiniGrid-DoorKey-16x16-v0 262144 20,2,1 24,2,1 28,2,1
sweep MiniGrid-MultiRoom-N6-v0 262144 20,2,1 28,2,1
