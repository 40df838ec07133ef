#!/bin/bash
# Round 2, fifteenth GPU call: launch-plan sweeps on the current kernels, e2e with calibrated stores and unequal chunks.
tag=${1:-r02o}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
echo "--- DoorKey plans (warps,vis,nbuf)"
for cfg in auto 10,2,2 12,2,2 14,2,2 16,2,2 18,2,2 19,2,2 19,1,2 20,1,2 24,2,1 28,2,1 32,1,1; do
  if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
  echo "$cfg: $($B 2>/dev/null | line)"
done
echo "--- FourRooms plans"
for cfg in auto 12,2,1 14,2,1 16,2,1 19,2,1 20,2,1 14,1,1 19,1,1 20,1,1; do
  if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
  echo "$cfg: $($B --env MiniGrid-FourRooms-v0 2>/dev/null | line)"
done
echo "--- Lava plans"
for cfg in auto 16,2,1 20,2,1 24,2,1 28,2,1 32,2,1 32,1,1 24,1,1; do
  if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
  echo "$cfg: $($B --env MiniGrid-LavaCrossingS9N1-v0 2>/dev/null | line)"
done
unset MINIGRID_B200_CFG
echo "--- Empty-8x8 x 65536 plans"
for cfg in auto 8,0,2 10,0,2 14,0,2 20,0,2 16,0,1 32,0,1; do
  if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
  echo "$cfg: $($B --env MiniGrid-Empty-8x8-v0 --envs-per-gpu 65536 2>/dev/null | line)"
done
unset MINIGRID_B200_CFG
echo "hotfirst=0: $(MINIGRID_B200_HOTFIRST=0 $B 2>/dev/null | line)"
echo "pdl=0: $(MINIGRID_B200_PDL=0 $B 2>/dev/null | line)"
echo "--- e2e"
E="timeout 200 python bench.py --no-cpu-baseline --no-configs --steps 200 --warmup 10"
e2e() { python -c "import json,sys;d=json.loads(sys.stdin.read());print(['%.3g'%v for v in d['e2e']['repetitions']], 'full', ['%.3g'%v for v in d['e2e']['full_format']['repetitions']])" 2>&1 | tail -1; }
echo "R=4 auto : $(MINIGRID_B200_HOST_TRACE=1 $E 2>$out/e2e_$tag.err | e2e)"; grep "host expansion" $out/e2e_$tag.err | head -4; tail -2 $out/e2e_$tag.err
echo "R=1 auto : $(MINIGRID_B200_HOST_TRACE=1 $E --rotate 1 2>$out/e2e1_$tag.err | e2e)"; grep "host expansion" $out/e2e1_$tag.err | head -2; tail -1 $out/e2e1_$tag.err
echo "R=4 chunks=12: $(MINIGRID_B200_HOST_CHUNKS=12 $E 2>/dev/null | e2e)"
echo "R=4 chunks=16: $(MINIGRID_B200_HOST_CHUNKS=16 $E 2>/dev/null | e2e)"
