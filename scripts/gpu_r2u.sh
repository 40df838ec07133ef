#!/bin/bash
# Round 2: two CTAs per SM for the one-buffer tiled kernels across kinds; the driver-style bench line with one graph per short run.
tag=${1:-r02u}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
sweep() { env=$1; n=$2; shift 2
  for cfg in "$@"; do
    if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
    echo "$env x $n $cfg: $(MINIGRID_B200_VERBOSE=1 $B --env $env --envs-per-gpu $n 2>$out/v.err | line) $(grep -m1 'K1 plan' $out/v.err | cut -c30-100)"
  done; unset MINIGRID_B200_CFG; }
sweep MiniGrid-DoorKey-8x8-v0 262144 11,2,1 12,2,1
sweep MiniGrid-Empty-8x8-v0 262144 auto 11,0,1
sweep MiniGrid-Empty-8x8-v0 65536 auto 11,0,1 7,0,1
sweep MiniGrid-LavaCrossingS9N1-v0 262144 auto 9,2,1 10,1,1
sweep MiniGrid-GoToDoor-8x8-v0 262144 auto 11,0,1
sweep MiniGrid-Fetch-8x8-N3-v0 262144 auto 11,0,1
sweep MiniGrid-DoorKey-5x5-v0 262144 auto 11,2,1
sweep MiniGrid-LavaGapS7-v0 262144 auto 11,2,1
sweep MiniGrid-FourRooms-v0 262144 auto 10,2,1 10,1,1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $out/${tag}_bench_k20.json 2> $out/${tag}_bench_k20.err; echo "k20 rc=$?"
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_k20.json"))
print(d["value"], d["roofline"]["frac"], d["run"]["launch"], d["e2e"]["value"], d["e2e"]["repetitions"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])])
PY
