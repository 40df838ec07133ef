"""Debug aid: FourRooms at full size against the oracle, reporting WHICH envs differ (tile, lane, position in its CTA's range)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from engine_adapter import make_engine
from oracle.oracle import OracleVecEnv
env_id = sys.argv[1] if len(sys.argv) > 1 else "MiniGrid-FourRooms-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
eng = make_engine(env_id, n, "next_step")
orc = OracleVecEnv(env_id, n, autoreset="next_step", n_threads=0)
eo, ed = eng.reset(seed=0); oo, od = orc.reset(seed=0)
print("reset equal:", np.array_equal(eo, oo))
rng = np.random.default_rng(1)
T = (n + 31) // 32
G = 148
tq, tr = T // G, T % G
for t in range(steps):
    a = rng.integers(0, 7, n).astype(np.int32)
    e = eng.step(a); o = orc.step(a)
    bad = np.nonzero((e[0] != o[0]).reshape(n, -1).any(1))[0]
    print(f"t={t}: {len(bad)} envs differ; dir diff {int((e[1] != o[1]).sum())}; term diff {int((e[3].astype(bool) != o[3]).sum())}")
    if len(bad):
        tiles = np.unique(bad // 32)
        # position of a tile in its CTA's range
        cta = np.zeros(len(tiles), int); pos = np.zeros(len(tiles), int)
        for i, tl in enumerate(tiles):
            c = 0
            lo = 0
            # CTA c owns [c*tq + min(c,tr), ...)
            c = min(int(tl // (tq + 1)), tr) if tr and tl < tr * (tq + 1) else (tr + (tl - tr * (tq + 1)) // tq if tq else 0)
            lo = c * tq + min(c, tr)
            cta[i], pos[i] = c, tl - lo
        print("   tiles:", len(tiles), "positions in CTA range (hist of pos//20):", np.bincount(pos // 20, minlength=4)[:6], "lanes hist:", np.bincount(bad % 32, minlength=32))
        print("   first bad envs:", bad[:10], "pos:", pos[:10])
        k = bad[0]
        print("   engine obs type plane:\n", e[0][k][:, :, 0].T, "\n   oracle:\n", o[0][k][:, :, 0].T)
        break
