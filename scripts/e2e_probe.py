"""End-to-end host path (mg_step_host): ms per 262144-env DoorKey step for the full and the packed format, against the
number of expander threads and copy chunks; MINIGRID_B200_HOST_TRACE prints where the time goes."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch

    from minigrid_b200 import MinigridVecEnv, bind_to_gpu_numa_node

    bind_to_gpu_numa_node(torch.device("cuda", 0))
    fmt, threads = sys.argv[2], int(sys.argv[3])
    n = 262144
    envs = [MinigridVecEnv("MiniGrid-DoorKey-8x8-v0", n) for _ in range(2)]
    for e in envs:
        e.reset(seed=0)
        e.set_host_format(fmt, threads)
    acts = torch.randint(0, 7, (16, n), dtype=torch.int32).pin_memory()
    for t in range(10):
        envs[t % 2].step_host(acts[t % 16])
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for t in range(60):
            envs[t % 2].step_host(acts[t % 16])
        best = min(best, (time.perf_counter() - t0) / 60)
    print(f"{fmt} threads={threads} chunks={os.environ.get('MINIGRID_B200_HOST_CHUNKS', 'auto')}: {best * 1e3:.3f} ms/step = {n / best / 1e6:.0f} M env-steps/s")
    for e in envs:
        e.close()
else:
    for fmt, threads, chunks in (("full", 0, None), ("packed", 8, None), ("packed", 12, None), ("packed", 15, None), ("packed", 16, None),
                                 ("packed", 15, 4), ("packed", 15, 2), ("packed", 15, 16)):
        env = dict(os.environ, MINIGRID_B200_HOST_TRACE="1")
        if chunks:
            env["MINIGRID_B200_HOST_CHUNKS"] = str(chunks)
        r = subprocess.run([sys.executable, __file__, "child", fmt, str(threads)], capture_output=True, text=True, env=env)
        print(r.stdout.strip(), "|", r.stderr.strip().splitlines()[-1][:260] if r.stderr.strip() else "")
