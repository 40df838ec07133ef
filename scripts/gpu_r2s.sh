#!/bin/bash
# Round 2, final 2-GPU call: the torchrun path of bench.py as the driver launches it (both arms, --steps 20 --warmup 3), on the final source.
tag=${1:-r02s}
out=gpurun_out
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > $out/${tag}_bench_ref_n2.json 2> $out/${tag}_bench_ref_n2.err; echo "ref N=2 rc=$?"; tail -1 $out/${tag}_bench_ref_n2.json | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err; echo "N=2 rc=$?"; tail -2 $out/${tag}_bench_n2.err
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench_n2.json").read().strip().splitlines()[-1])
print("N=2", d["value"], d["n_gpus"], d["roofline"]["frac"], [(c["env"], c["total_envs"], round(c["value"] / 1e9, 2)) for c in d.get("configs", [])], d["e2e"], d["run"]["numa"])
PY
