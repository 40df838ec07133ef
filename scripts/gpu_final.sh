#!/bin/bash
# One gpurun call for the round's evidence: GPU parity tests, the default bench line, the reference arm, ncu launch
# list + full capture of k_step, bench lines of the other generators, and the step timeline (debug build).
# usage: scripts/gpu_final.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
timeout 300 python bench.py > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_$tag.log | cut -c1-2500
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_$tag.log | cut -c1-600
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -o gpurun_out/prof_kstep_$tag python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1; echo "ncu full rc=$?"
for env in MiniGrid-Empty-8x8-v0 MiniGrid-LavaCrossingS9N1-v0 MiniGrid-FourRooms-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-MultiRoom-N6-v0; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_${tag}_$env.log 2>/dev/null
  echo "$env: $(tail -1 gpurun_out/bench_${tag}_$env.log | cut -c1-120)"
done
if [ -f minigrid_b200/libminigrid_b200_tl.so ]; then
  MINIGRID_B200_LIB=$PWD/minigrid_b200/libminigrid_b200_tl.so timeout 60 python scripts/timeline.py 262144 > gpurun_out/timeline_$tag.txt 2>&1; tail -3 gpurun_out/timeline_$tag.txt
fi
