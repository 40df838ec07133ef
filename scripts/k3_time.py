"""K3 (k_full_obs) time and roofline fraction: FourRooms / DoorKey x 262144."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minigrid_b200 import MinigridVecEnv
for env_id in ("MiniGrid-FourRooms-v0", "MiniGrid-DoorKey-8x8-v0", "MiniGrid-DoorKey-16x16-v0"):
    n = 262144
    es = [MinigridVecEnv(env_id, n) for _ in range(2)]
    outs = []
    for e in es:
        e.reset(seed=0)
        outs.append(torch.empty((n, e.width, e.height, 3), dtype=torch.uint8, device="cuda"))
    for i in range(4): es[i % 2].full_obs(outs[i % 2])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40): es[i % 2].full_obs(outs[i % 2])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 40
    e = es[0]
    by = n * (e.width * e.height * 4 + 16)
    print(f"{env_id}: {ms*1e3:.1f} us, {by/ms/1e6:.0f} GB/s algorithmic, frac {by/ms/1e6/6576.4:.3f}")
