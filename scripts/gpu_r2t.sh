#!/bin/bash
# Round 2: the bench line as the driver runs it (--steps 20 --warmup 3) with the e2e leg decoupled from --steps; two CTAs per SM.
tag=${1:-r02t}
out=gpurun_out
mkdir -p $out
B="timeout 150 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 --steps 800 --warmup 20"
line() { python -c "import json,sys;d=json.loads(sys.stdin.read());print('%.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'])" 2>&1 | tail -1; }
for cfg in auto 11,2,1 10,2,1 7,2,1 22,2,1; do
  if [ $cfg = auto ]; then unset MINIGRID_B200_CFG; else export MINIGRID_B200_CFG=$cfg; fi
  echo "DoorKey $cfg: $(MINIGRID_B200_VERBOSE=1 $B 2>$out/v.err | line) $(grep -m1 'K1 plan' $out/v.err | cut -c1-120)"
done
unset MINIGRID_B200_CFG
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $out/${tag}_bench_k20.json 2> $out/${tag}_bench_k20.err; echo "k20 rc=$?"
python - <<PY
import json
d = json.load(open("$out/${tag}_bench_k20.json"))
print(d["value"], d["roofline"]["frac"], d["e2e"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d["cpu_baseline"]["value"])
PY
