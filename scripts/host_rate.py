import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minigrid_b200 import MinigridVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
e = MinigridVecEnv("MiniGrid-DoorKey-8x8-v0", n); e.reset(seed=0)
acts = torch.randint(0, 7, (64, n), device="cuda", dtype=torch.int32)
rows = [acts[i] for i in range(64)]
for t in range(50): e.step(rows[t % 64])
torch.cuda.synchronize()
T = 3000
t0 = time.perf_counter()
for t in range(T): e.step(rows[t % 64])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"n={n}: host enqueue {1e6*(t1-t0)/T:.2f} us/step, total {1e6*(t2-t0)/T:.2f} us/step")
