#!/bin/bash
# After the racecheck findings of r02_final (a missing __syncwarp at warp_reset's entry for SAME_STEP, K3's barrier polled by every
# thread): parity suite, sanitizer and bench lines of the final source.
tag=${1:-r02_final2}
out=gpurun_out
mkdir -p $out
for tool in racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=2mg --print-limit 30 python scripts/sanitize_smoke.py > $out/${tag}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all cases' $out/${tag}_sanitizer_${tool}.log | tr '\n' ' ')"
done
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; tail -c 600 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 > $out/${tag}_bench_k20.json 2>/dev/null; echo "k20: $(cut -c1-160 $out/${tag}_bench_k20.json)"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > $out/${tag}_bench_reference_arm.json 2>&1; echo "ref: $(cut -c1-160 $out/${tag}_bench_reference_arm.json)"
timeout 120 python scripts/k3_time.py 2>&1 | tail -3
timeout 200 ncu --set full --clock-control none -k regex:k_full_obs -c 2 -o $out/${tag}_prof_fullobs python -c "
import torch
from minigrid_b200 import MinigridVecEnv
e = MinigridVecEnv('MiniGrid-FourRooms-v0', 262144); e.reset(seed=0)
for _ in range(3): e.full_obs()
torch.cuda.synchronize()" > /dev/null 2>&1; echo "full k3 rc=$?"
