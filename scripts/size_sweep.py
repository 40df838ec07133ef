"""K1 time per step against the number of environments (single batch, CUDA-graph replay, events): separates a
per-launch cost from a per-tile cost when comparing two builds (MINIGRID_B200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minigrid_b200 import MinigridVecEnv

env_id = sys.argv[1] if len(sys.argv) > 1 else "MiniGrid-DoorKey-8x8-v0"
sizes = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4736, 32768, 131072, 262144, 524288, 1048576]
out = []
for n in sizes:
    e = MinigridVecEnv(env_id, n, autoreset_mode=os.environ.get("SWEEP_AUTORESET", "next_step")); e.reset(seed=0)
    acts = torch.randint(0, 7, (64, n), device="cuda", dtype=torch.int32)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for t in range(20): e.step(acts[t % 64])
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for t in range(64): e.step(acts[t])
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for rep in range(5):
        a.record()
        for _ in range(10): g.replay()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1000 / 640)
    out.append(f"{n}:{best:.2f}")
    del e, g
print(" ".join(out))
