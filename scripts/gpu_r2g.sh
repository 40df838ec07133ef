#!/bin/bash
# Round 2, seventh GPU call (2 GPUs): the torchrun path of bench.py (both arms), then single-GPU checks of the RNG prefetch
# and the lower-latency host pool.
tag=${1:-r02g}
out=gpurun_out
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $out/bench_${tag}_ref2.json 2> $out/bench_${tag}_ref2.err; echo "ref N=2 rc=$?"; tail -1 $out/bench_${tag}_ref2.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > $out/bench_${tag}_n2.json 2> $out/bench_${tag}_n2.err; echo "N=2 rc=$?"; tail -2 $out/bench_${tag}_n2.err
python - <<PY
import json
d = json.loads(open("$out/bench_${tag}_n2.json").read().strip().splitlines()[-1])
print("N=2", d["value"], d["n_gpus"], d["roofline"]["frac"], [(c["env"], c["total_envs"], round(c["value"] / 1e9, 2)) for c in d.get("configs", [])], d["e2e"], d["run"]["numa"])
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench_${tag}_n1.json 2> $out/bench_${tag}_n1.err; echo "N=1 K=20 rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_${tag}_n1.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["roofline"]["frac"], [(c["env"], round(c["value"] / 1e9, 2), round(c["frac"], 3)) for c in d.get("configs", [])], d.get("autoreset_cost"), d["e2e"], d.get("full_obs"), d.get("config1"), d["cpu_baseline"]["value"])
PY
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --no-cpu-baseline --no-configs > $out/bench_${tag}_n1_1000.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("$out/bench_${tag}_n1_1000.json").read().strip().splitlines()[-1])
print("N=1 K=1000", d["value"], d["roofline"]["frac"], d["e2e"])
PY
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "packed or full_tiles or lockstep" 2>&1 | tail -2
