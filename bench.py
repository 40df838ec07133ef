#!/usr/bin/env python
"""bench.py — env-steps/s of the batched Minigrid hot path (BASELINE.json's metric).

    python bench.py --gpus 1 --steps K --warmup W            # this engine, 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # N GPUs, weak scaling
    python bench.py --impl reference ...                      # the reference algorithm on the host CPUs

One "step" = one lockstep vector step (action -> state', obs, reward, terminated, truncated, autoreset) of the
whole batch. Workload (config.workload): BASELINE.json configs[2], MiniGrid-DoorKey-8x8-v0 with 262144
environments per GPU (the configuration the >=1e8 steps/s target is quoted on), uniform random actions generated
on the device before timing, NEXT_STEP autoreset. Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_STEP = 348  # action 4 + patch 147 + obs 147 + reward 8 + flags 2 + agent 20 r + 20 w
L2_BYTES = 126e6
# dram__bytes_read.sum + dram__bytes_write.sum of one k_step launch from the committed ncu --set full capture
# (profiles/r01_final_kstep_summary.txt: 47.27 MB read + 13.2-13.3 MB written; obs writes mostly stay in L2)
TRAFFIC_BYTES_PER_LAUNCH = 60.5e6
TRAFFIC_SOURCE = "ncu --set full, profiles/r01_final_kstep_summary.txt (DoorKey-8x8 x 262144; applies to the default workload only)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--env", default="MiniGrid-DoorKey-8x8-v0")
    ap.add_argument("--envs-per-gpu", type=int, default=262144)
    ap.add_argument("--rotate", type=int, default=0, help="independent env batches cycled through so the working set exceeds L2 (0 = auto)")
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--graph", type=int, default=1, help="replay the step loop as CUDA graphs of this many steps x rotate (0 = eager launches)")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML from a background thread DURING the timed region."""

    def __init__(self, gpu_index, period_s=0.02):
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = None
        self._thread = None

    def start(self):
        import threading

        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return
        names = {
            pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                self._stop.wait(self.period)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
        return out


def usable_cores():
    """Host threads this process can actually run at once: min(online CPUs, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(np.ceil(int(quota) / int(period)))))
    except Exception:
        pass
    return n


def run_reference(args, rank, world):
    """The reference algorithm on the host CPUs: the oracle port (C restatement of MiniGridEnv.step/gen_obs,
    validated against the Python reference), one env slice per host thread. The Python reference itself cannot
    travel to the GPU box (no gymnasium in the image)."""
    if rank != 0:
        return
    from oracle.oracle import OracleVecEnv, max_threads

    n = args.envs_per_gpu
    cores = min(max_threads(), usable_cores())
    env = OracleVecEnv(args.env, n, autoreset="next_step", n_threads=cores)
    env.reset(seed=0)
    rng = np.random.default_rng(1234)
    # bounded sample: keep the whole run to a few minutes whatever the core count
    probe_steps = 2
    secs, _ = env.rollout(rng.integers(0, 7, (probe_steps, n)).astype(np.int32), n_threads=cores)
    per_step = max(secs / probe_steps, 1e-6)
    budget = 120.0
    steps = int(max(3, min(args.steps, budget / per_step)))
    warm = int(max(1, min(args.warmup, 10.0 / per_step)))
    env.rollout(rng.integers(0, 7, (warm, n)).astype(np.int32), n_threads=cores)
    secs, _ = env.rollout(rng.integers(0, 7, (steps, n)).astype(np.int32), n_threads=cores)
    value = n * steps / secs
    line = {
        "impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.env}, {n} envs, uniform random actions, NEXT_STEP autoreset", "env": args.env,
                   "envs": n, "host_threads": cores},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "per_core": value / cores, "kind": "port",
                         "sample": f"{n} envs x {steps} lockstep steps, C port of the reference algorithm (oracle/mg_oracle.c), {cores} threads"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(args):
    from oracle.oracle import OracleVecEnv, max_threads

    cores = min(max_threads(), usable_cores())
    n = 65536
    env = OracleVecEnv(args.env, n, autoreset="next_step", n_threads=cores)
    env.reset(seed=0)
    rng = np.random.default_rng(1234)
    secs, _ = env.rollout(rng.integers(0, 7, (2, n)).astype(np.int32), n_threads=cores)
    steps = int(max(4, min(2000, args.cpu_seconds / max(secs / 2, 1e-6))))
    secs, _ = env.rollout(rng.integers(0, 7, (steps, n)).astype(np.int32), n_threads=cores)
    return {"value": n * steps / secs, "unit": "env-steps/s", "cores": cores, "per_core": n * steps / secs / cores, "kind": "port",
            "sample": f"{n} envs x {steps} lockstep steps of {args.env} ({secs:.1f} s), oracle C port on {cores} host threads "
                      f"(os.cpu_count()={os.cpu_count()}, cgroup CPU quota respected)"}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from minigrid_b200 import make_sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.envs_per_gpu
    total = n * world
    K, W = args.steps, args.warmup
    # working set of one batch: grid tiles + obs + agent/rng/outputs; rotate enough batches to exceed L2
    probe = make_sharded(args.env, total, rank, world, device=dev)
    wpe_bytes = ((probe.height + 2) * ((probe.width + 3) // 4) + (probe.width + 2) * ((probe.height + 3) // 4)) * 4
    ws = n * (wpe_bytes + 147 + 16 + 4 + 4 + 8 + 2)
    R = args.rotate if args.rotate > 0 else max(1, int(np.ceil(2.2 * L2_BYTES / ws)))
    batches = [probe] + [make_sharded(args.env, total, rank, world, device=dev) for _ in range(R - 1)]
    for b, e in enumerate(batches):
        e.reset(seed=1_000_003 * b)  # env i of batch b: seed 1000003*b + global index
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    T = min(K + W, 512)  # action table rows, cycled
    actions = torch.randint(0, 7, (T, n), generator=gen, device=dev, dtype=torch.int32)
    torch.cuda.synchronize()

    act_rows = [actions[i] for i in range(T)]  # views made once: the timed loop is launches only
    step_fns = [b.step for b in batches]

    def run(steps, first=0):
        for t in range(first, first + steps):
            step_fns[t % R](act_rows[t % T])

    run(W)
    if W < 3:
        run(3 - W, W)  # never fewer than 3 untimed steps before the timed region, whatever --warmup says
    barrier()
    # The step loop is launch-bound for small batches (one ~20 us kernel per step vs ~8 us of Python + driver per
    # launch), so it is captured once as a CUDA graph through the same public step() calls and replayed. K steps
    # are still exactly K kernel launches on the device.
    graph = None
    G = 0
    if args.graph:
        G = R * max(1, min(T // R, 128 // R))  # steps per graph: every batch and a run of distinct action rows
        if G > K:
            G = 0
    if G:
        try:
            cap_stream = torch.cuda.Stream(device=dev)
            cap_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cap_stream):
                run(G)  # warm the capture stream
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=cap_stream):
                    run(G)
            torch.cuda.current_stream(dev).wait_stream(cap_stream)
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001  (fall back to eager launches, and say so)
            graph = None
            G = 0
            graph_error = repr(exc)
            torch.cuda.synchronize()
    eager_run = run
    if graph is not None:
        def run(steps, first=0):  # noqa: F811
            for _ in range(steps // G):
                graph.replay()
            if steps % G:
                eager_run(steps % G, first)
        run(G)
        barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = sum(b.launch_count for b in batches)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    h0 = time.perf_counter()
    run(K, W)
    host_enqueue_s = time.perf_counter() - h0
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = sum(b.launch_count for b in batches) - l0
    if graph is not None:
        launches += (K // G) * G  # launches replayed by the graphs (each captured step() is one kernel node)
    for b in batches:
        b.check_actions()
    value = total * K / (ms * 1e-3)

    # the same loop with a single L2-resident batch, for context (not the headline)
    run(W)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):  # eager launches, one batch: shows the launch-bound regime next to the graph number
        step_fns[0](act_rows[t % T])
    e1.record()
    torch.cuda.synchronize()
    ms_res = max_over_ranks(e0.elapsed_time(e1))

    # dominant kernel: a vector step is exactly ONE K1 launch (launches == K), so the CUDA events
    # that bracket the timed region measure K back-to-back k_step launches on the launching stream; ms / K is the
    # average launch duration including launch gaps (conservative)
    kstep_ms = ms / K if launches == K else None
    # cross-check with per-launch events (perturbs the stream; reported, not used for the headline)
    kstep_ms_events = None
    try:
        for b in batches:
            b.profile_kernels(True)
        eager_run(min(K, 200), W)  # graph replays bypass the C-ABI call that records the events
        torch.cuda.synchronize()
        tot, cnt = 0.0, 0
        for b in batches:
            t_ms, c = b.kernel_time_ms()
            tot += t_ms; cnt += c
            b.profile_kernels(False)
        kstep_ms_events = tot / max(cnt, 1)
    except AttributeError:
        pass

    # end to end through the host-buffer API: pinned host actions in, pinned host obs/reward/flags out
    Ke = min(args.e2e_steps, K)
    host_actions = torch.randint(0, 7, (min(Ke, 64), n), dtype=torch.int32).pin_memory()
    for t in range(max(3, 2 * R)):  # every batch allocates its pinned staging on first use: keep that out of the timing
        batches[t % R].step_host(host_actions[t % host_actions.shape[0]])
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        batches[t % R].step_host(host_actions[t % host_actions.shape[0]])
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total * Ke / e2e_s
    h2d = n * 4
    d2h = n * (147 + 4 + 8 + 1 + 1)

    peak, peak_src = load_peaks()
    default_workload = args.env == "MiniGrid-DoorKey-8x8-v0" and n == 262144  # what the committed ncu capture measured
    roof = None
    if kstep_ms:
        achieved = ALGO_BYTES_PER_STEP * n / (kstep_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "k_step (K1: transition + gen_obs)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": TRAFFIC_BYTES_PER_LAUNCH if default_workload else None,
                "traffic_source": TRAFFIC_SOURCE if default_workload else None,
                "peak_source": peak_src, "kernel_ms": kstep_ms, "kernel_ms_per_launch_events": kstep_ms_events,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_STEP * n}

    if rank == 0:
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": f"{args.env}, {n} envs per GPU ({total} total), uniform random actions, NEXT_STEP autoreset",
                       "env": args.env, "envs_per_gpu": n, "total_envs": total, "autoreset": "next_step",
                       "l2": f"{R} independent env batches cycled, working set {R * ws / 1e6:.0f} MB per GPU > 126 MB L2 (inputs larger than L2)",
                       "parallelism": f"env-sharded x{world}, no collective on the step path",
                       "launch": (f"CUDA graph replay, {G} steps per graph" if graph is not None else "eager launches")},
            "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke},
            "gpu_launches": int(launches),
            "value_l2_resident": total * K / (ms_res * 1e-3),
            "host_enqueue_us_per_step": 1e6 * host_enqueue_s / K,
        }
        if roof:
            line["roofline"] = roof
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
