#!/bin/bash
# Round 2, second GPU call: register-window gather, hot-tile-first order, packed host format.
tag=${1:-r02b}
out=gpurun_out
mkdir -p $out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q --timeout 120 > $out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 $out/pytest_$tag.log
echo "--- size sweeps (us per step, one batch, graph replay, 3400 steps: waves included)"
for arm in cur cur_nohot; do
  L=$PWD/minigrid_b200/libminigrid_b200.so; [ $arm = base ] && L=$PWD/minigrid_b200/libminigrid_b200_base.so
  H=1; [ $arm = cur_nohot ] && H=0
  for env in MiniGrid-DoorKey-8x8-v0 MiniGrid-LavaCrossingS9N1-v0 MiniGrid-FourRooms-v0 MiniGrid-DoorKey-16x16-v0 MiniGrid-MultiRoom-N6-v0; do
    echo "$arm $env: $(MINIGRID_B200_HOTFIRST=$H MINIGRID_B200_LIB=$L timeout 100 python scripts/size_sweep.py $env 262144 2>&1 | tail -1)"
  done
done
echo "--- bench (desynchronised) hot-first on / off"
timeout 600 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; echo "bench rc=$?"; tail -c 3500 $out/bench_$tag.json; tail -3 $out/bench_$tag.err
MINIGRID_B200_HOTFIRST=0 timeout 300 python bench.py --no-cpu-baseline > $out/bench_${tag}_nohot.json 2>/dev/null; echo "nohot: $(cut -c1-150 $out/bench_${tag}_nohot.json)"
python - <<PY
import json
for f in ("gpurun_out/bench_TAG.json", "gpurun_out/bench_TAG_nohot.json"):
    try:
        d = json.load(open(f.replace("TAG", "${tag}")))
        print(f, d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"])
    except Exception as e:
        print(f, e)
PY
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > $out/bench_${tag}_k20.json 2>/dev/null; echo "k20: $(cut -c1-150 $out/bench_${tag}_k20.json)"
for env in MiniGrid-DoorKey-16x16-v0 MiniGrid-MultiRoom-N6-v0 MiniGrid-LockedRoom-v0 MiniGrid-Playground-v0 MiniGrid-GoToDoor-8x8-v0 MiniGrid-Fetch-8x8-N3-v0 MiniGrid-PutNear-8x8-N3-v0 MiniGrid-RedBlueDoors-8x8-v0 MiniGrid-MemoryS13Random-v0 MiniGrid-GoToObject-8x8-N2-v0; do
  timeout 120 python bench.py --env $env --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 10 > $out/bench_${tag}_$env.json 2>/dev/null
  echo "$env: $(python -c "import json;d=json.load(open('$out/bench_${tag}_$env.json'));print(d['value'], d['roofline']['frac'], d['e2e']['value'])")"
done
echo "--- sanitizer"
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=2mg --print-limit 30 python scripts/sanitize_smoke.py > $out/sanitizer_${tool}_$tag.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all cases' $out/sanitizer_${tool}_$tag.log | tr '\n' ' ')"
done
echo "--- ncu"
for cfg in MiniGrid-DoorKey-8x8-v0:262144 MiniGrid-Empty-8x8-v0:65536 MiniGrid-LavaCrossingS9N1-v0:262144 MiniGrid-FourRooms-v0:262144; do
  env=${cfg%%:*}; n=${cfg##*:}
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:k_step -s 24 -c 24 --csv --log-file $out/${tag}_traffic_$env.csv \
    python bench.py --env $env --envs-per-gpu $n --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1
  echo "traffic $env rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/prof_${tag}_doorkey \
  python bench.py --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full doorkey rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/prof_${tag}_fourrooms \
  python bench.py --env MiniGrid-FourRooms-v0 --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full fourrooms rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o $out/prof_${tag}_lava \
  python bench.py --env MiniGrid-LavaCrossingS9N1-v0 --steps 60 --warmup 4 --graph 0 --no-cpu-baseline --no-configs --e2e-steps 1 > /dev/null 2>&1; echo "full lava rc=$?"
