#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, and ncu evidence (launch list + full capture of k_step).
# usage: scripts/gpu_check.sh <tag> [bench args...]
tag=${1:-run}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
timeout 400 python bench.py --steps 1000 --warmup 50 "$@" > gpurun_out/bench_$tag.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_$tag.log | cut -c1-3000
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 "$@" > /dev/null 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 2 -o gpurun_out/prof_kstep_$tag python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 "$@" > /dev/null 2>&1; echo "ncu full rc=$?"
