#!/bin/bash
# Round 2, a short call: racecheck of K3 after re-arming its barrier; the window kernel at 24 warps / 80 registers (A/B library).
tag=${1:-r02r}
out=gpurun_out
mkdir -p $out
timeout 600 compute-sanitizer --tool racecheck --kernel-regex kns=2mg --print-limit 30 python scripts/sanitize_smoke.py 4 > $out/${tag}_sanitizer_racecheck_k3.log 2>&1
echo "racecheck rc=$? $(grep -E 'RACECHECK SUMMARY|all cases' $out/${tag}_sanitizer_racecheck_k3.log | tr '\n' ' ')"
timeout 300 python -m pytest tests -m gpu -x -q --timeout 300 -k "wrapper_classes or shards or (lockstep_vs_oracle and (Empty-5x5 or FourRooms))" 2>&1 | tail -2
timeout 120 python scripts/k3_time.py 2>&1 | tail -3
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
for env in MiniGrid-FourRooms-v0 MiniGrid-DoorKey-16x16-v0; do
  echo "$env cur (20 warps, 96 regs): $($B --env $env 2>/dev/null | line)"
  for cfg in 24,2,1 22,2,1 24,1,1; do echo "$env w24 lib cfg=$cfg: $(MINIGRID_B200_CFG=$cfg MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_w24.so $B --env $env 2>/dev/null | line)"; done
done
