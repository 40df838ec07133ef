import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minigrid_b200 import MinigridVecEnv
n = 262144
e = MinigridVecEnv("MiniGrid-DoorKey-8x8-v0", n)
e.reset(seed=0)
acts = torch.randint(0, 7, (8, n), dtype=torch.int32).pin_memory()
for t in range(5): e.step_host(acts[t % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(50): e.step_host(acts[t % 8])
dt = (time.perf_counter() - t0) / 50
print("step_host ms", dt * 1e3, "steps/s", n / dt)
# components
hb = e._host_buffers()
d_obs = torch.empty((n, 7, 7, 3), dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for name, host, dev in [("obs 38.5MB", hb["image"], d_obs)]:
        for _ in range(3): host.copy_(dev, non_blocking=True); s.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): host.copy_(dev, non_blocking=True); s.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print(name, "sync copy ms", dt * 1e3, "GB/s", host.numel() / dt / 1e9)
print("numa/affinity", os.sched_getaffinity(0).__len__())
os.system(r"numactl --show 2>/dev/null | head -3; cat /proc/self/status | grep -i 'Mems_allowed_list\|Cpus_allowed_list'")
