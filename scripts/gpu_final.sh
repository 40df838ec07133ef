#!/bin/bash
# One gpurun call for the round's final evidence: GPU parity tests, the default bench line (as the driver runs it and with
# the default step count), the reference arm, ncu launch list + full capture of k_step, DRAM traffic of the BASELINE
# configs, compute-sanitizer, and a bench line per widened env kind.
# usage: scripts/gpu_final.sh <tag>
tag=${1:-r02_final}
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -c 1500 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > $out/${tag}_bench_k20.json 2>/dev/null; echo "k20: $(cut -c1-160 $out/${tag}_bench_k20.json)"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $out/${tag}_bench_reference_arm.json 2>&1; echo "ref: $(cut -c1-160 $out/${tag}_bench_reference_arm.json)"
echo "--- ncu"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $out/${tag}_launches.csv python bench.py --steps 60 --warmup 5 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 3 > /dev/null 2>&1; echo "launch list rc=$?"
for cfg in MiniGrid-DoorKey-8x8-v0:262144 MiniGrid-Empty-8x8-v0:65536 MiniGrid-LavaCrossingS9N1-v0:262144 MiniGrid-FourRooms-v0:262144; do
  env=${cfg%%:*}; n=${cfg##*:}
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:k_step -s 24 -c 24 --csv --log-file $out/${tag}_traffic_$env.csv \
    python bench.py --env $env --envs-per-gpu $n --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1
  echo "traffic $env rc=$?"
done
for w in doorkey:MiniGrid-DoorKey-8x8-v0 lava:MiniGrid-LavaCrossingS9N1-v0 fourrooms:MiniGrid-FourRooms-v0; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/${tag}_prof_${w%%:*} \
    python bench.py --env ${w##*:} --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full ${w%%:*} rc=$?"
done
timeout 200 ncu --set full --clock-control none -k regex:k_full_obs -c 2 -o $out/${tag}_prof_fullobs python -c "
import torch
from minigrid_b200 import MinigridVecEnv
e = MinigridVecEnv('MiniGrid-FourRooms-v0', 262144); e.reset(seed=0)
for _ in range(3): e.full_obs()
torch.cuda.synchronize()" > /dev/null 2>&1; echo "full k3 rc=$?"
echo "--- sanitizer"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=2mg --print-limit 30 python scripts/sanitize_smoke.py > $out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all cases' $out/${tag}_sanitizer_${tool}.log | tr '\n' ' ')"
done
echo "--- one bench line per env family"
timeout 120 python scripts/k3_time.py 2>&1 | tail -3
for env in MiniGrid-Empty-8x8-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-MultiRoom-N6-v0 MiniGrid-LockedRoom-v0 MiniGrid-Playground-v0 MiniGrid-GoToDoor-8x8-v0 MiniGrid-Fetch-8x8-N3-v0 MiniGrid-PutNear-8x8-N3-v0 MiniGrid-RedBlueDoors-8x8-v0 MiniGrid-MemoryS13Random-v0 MiniGrid-GoToObject-8x8-N2-v0 MiniGrid-Dynamic-Obstacles-8x8-v0 MiniGrid-Unlock-v0 MiniGrid-BlockedUnlockPickup-v0 MiniGrid-KeyCorridorS6R3-v0 MiniGrid-ObstructedMaze-Full-v1 MiniGrid-LavaGapS7-v0 MiniGrid-DistShift2-v0; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 10 > $out/${tag}_bench_$env.json 2>/dev/null
  echo "$env: $(python -c "import json;d=json.load(open('$out/${tag}_bench_$env.json'));print('%.3g'%d['value'], '%.3f'%d['roofline']['frac'], 'autoreset/step %.4f'%d['run']['autoreset_fraction_per_step'], 'e2e %.3g'%d['e2e']['value'])" 2>&1 | tail -1)"
done
