"""Host-side expansion of the packed step records (mg_expand_packed_mt): milliseconds per 262144-env step against the
number of threads, with plain and with non-temporal stores (MINIGRID_B200_EXPAND_STREAM=1). CPU only."""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import torch

    from minigrid_b200 import _lib

    L = _lib.load()
    n = 262144
    rng = np.random.default_rng(0)
    packed = rng.integers(0, 256, (n, 52), dtype=np.uint8)
    packed[:, :49] &= 0x7F
    packed[:, 49] &= 0x0F
    pin = torch.cuda.is_available()
    mk = lambda shape, dt: (torch.zeros(shape, dtype=dt).pin_memory() if pin else torch.zeros(shape, dtype=dt))  # noqa: E731
    obs, d, r, te, tr = mk((n, 147), torch.uint8), mk(n, torch.int32), mk(n, torch.float64), mk(n, torch.uint8), mk(n, torch.uint8)
    src = mk((n, 52), torch.uint8)
    src.copy_(torch.from_numpy(packed))
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    out = []
    for T in [int(x) for x in sys.argv[2].split(",")]:
        best = 1e9
        for rep in range(8):
            t0 = time.perf_counter()
            L.mg_expand_packed_mt(p(src), n, 640, p(obs), p(d), p(r), p(te), p(tr), T)
            best = min(best, time.perf_counter() - t0)
        out.append(f"{T}:{best * 1e3:.3f}")
    print(" ".join(out), "(pinned)" if pin else "(pageable)")
else:
    threads = sys.argv[1] if len(sys.argv) > 1 else "1,2,4,8,12,15,16,24,32"
    for name, env in (("plain stores", {}), ("streaming stores", {"MINIGRID_B200_EXPAND_STREAM": "1"})):
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, __file__, "child", threads], capture_output=True, text=True, env=e)
        print(name, "threads:ms", r.stdout.strip(), r.stderr.strip()[-300:])
