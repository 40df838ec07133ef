"""Per-step GPU time from a fresh reset (CUDA events around every step): separates the steady state from the
synchronised truncation waves (DoorKey: every env truncates at step 640)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from minigrid_b200 import MinigridVecEnv
env_id = sys.argv[1] if len(sys.argv) > 1 else "MiniGrid-DoorKey-8x8-v0"
n, T = 262144, int(sys.argv[2]) if len(sys.argv) > 2 else 1400
e = MinigridVecEnv(env_id, n); e.reset(seed=0)
gen = torch.Generator(device="cuda").manual_seed(1)
acts = torch.randint(0, 7, (64, n), generator=gen, device="cuda", dtype=torch.int32)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(T + 1)]
for t in range(8): e.step(acts[t % 64])
e.reset(seed=0); torch.cuda.synchronize()
ev[0].record()
for t in range(T):
    e.step(acts[t % 64]); ev[t + 1].record()
torch.cuda.synchronize()
dt = np.array([ev[t].elapsed_time(ev[t + 1]) * 1e3 for t in range(T)])
print(env_id, "median %.2f us, mean %.2f us, p99 %.2f, max %.1f at step %d" % (np.median(dt), dt.mean(), np.percentile(dt, 99), dt.max(), dt.argmax() + 1))
big = np.where(dt > 3 * np.median(dt))[0]
print("slow steps:", [(int(i) + 1, round(float(dt[i]), 1)) for i in big[:12]])
