#!/bin/bash
# Round 2: bench.py after its last edit (expander threads per rank counted before the NUMA binding): a short run.
timeout 200 python bench.py --no-cpu-baseline --no-configs --steps 100 --warmup 5 2>&1 | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['host_threads'])"
