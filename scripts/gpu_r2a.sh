#!/bin/bash
# Round 2, first GPU call: parity of everything wired on the CPU, same-box A/B against the round-1 library, the new
# bench line, compute-sanitizer, and ncu captures (traffic per workload, full sets for the kernels VERDICT r1 names).
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_$tag.log
echo "--- A/B size sweeps (us per step, one batch, graph replay; waves included: 3400 steps)"
for lib in base cur; do
  L=$PWD/minigrid_b200/libminigrid_b200_$lib.so; [ $lib = cur ] && L=$PWD/minigrid_b200/libminigrid_b200.so
  for ar in next_step disabled; do
    echo "$lib $ar DoorKey: $(MINIGRID_B200_LIB=$L SWEEP_AUTORESET=$ar timeout 100 python scripts/size_sweep.py MiniGrid-DoorKey-8x8-v0 262144,1048576 2>&1 | tail -1)"
    echo "$lib $ar FourRooms: $(MINIGRID_B200_LIB=$L SWEEP_AUTORESET=$ar timeout 100 python scripts/size_sweep.py MiniGrid-FourRooms-v0 262144 2>&1 | tail -1)"
  done
  echo "$lib next_step Lava: $(MINIGRID_B200_LIB=$L timeout 100 python scripts/size_sweep.py MiniGrid-LavaCrossingS9N1-v0 262144 2>&1 | tail -1)"
done
echo "--- bench"
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -c 3000 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > $out/bench_${tag}_k20.json 2>/dev/null; echo "k20: $(cut -c1-200 $out/bench_${tag}_k20.json)"
timeout 120 python bench.py --sync-episodes --no-cpu-baseline --no-configs > $out/bench_${tag}_sync.json 2>/dev/null; echo "sync: $(cut -c1-200 $out/bench_${tag}_sync.json)"
echo "--- sanitizer"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=2mg --print-limit 30 python scripts/sanitize_smoke.py > $out/sanitizer_${tool}_$tag.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all cases' $out/sanitizer_${tool}_$tag.log | tr '\n' ' ')"
done
echo "--- ncu traffic"
for cfg in MiniGrid-DoorKey-8x8-v0:262144 MiniGrid-Empty-8x8-v0:65536 MiniGrid-LavaCrossingS9N1-v0:262144 MiniGrid-FourRooms-v0:262144; do
  env=${cfg%%:*}; n=${cfg##*:}
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:k_step -s 24 -c 24 --csv --log-file $out/${tag}_traffic_$env.csv \
    python bench.py --env $env --envs-per-gpu $n --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1
  echo "traffic $env rc=$?"
done
echo "--- ncu full"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/prof_${tag}_doorkey \
  python bench.py --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full doorkey rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/prof_${tag}_lava \
  python bench.py --env MiniGrid-LavaCrossingS9N1-v0 --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full lava rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/prof_${tag}_fourrooms \
  python bench.py --env MiniGrid-FourRooms-v0 --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full fourrooms rc=$?"
# a synchronised truncation wave: DoorKey step 640 of one batch (launch index: 3 warm-up + 640 ...)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 639 -c 3 -o $out/prof_${tag}_wave \
  python bench.py --sync-episodes --rotate 1 --steps 700 --warmup 3 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full wave rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $out/launches_$tag.csv python bench.py --steps 60 --warmup 5 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 3 > /dev/null 2>&1; echo "ncu list rc=$?"
ls -la $out | tail -30
