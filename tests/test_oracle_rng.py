"""numpy restatement inside the oracle (SeedSequence -> PCG64 -> integers/shuffle/next32) against numpy."""
import numpy as np

from oracle.oracle import OracleVecEnv


def _gen(seed):
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence(int(seed))))


def test_seed_sequence_and_state_words():
    seeds = np.array([0, 1, 2, 123, 2**31, 2**32 - 1, 2**32, 2**40 + 17, 2**63 + 5, 2**64 - 1], dtype=np.uint64)
    o = OracleVecEnv("MiniGrid-Empty-5x5-v0", len(seeds))
    o.seed(seeds)
    rng = o.get_state()["rng"]
    m = (1 << 64) - 1
    for i, s in enumerate(seeds):
        st = np.random.PCG64(np.random.SeedSequence(int(s))).state
        want = [st["state"]["state"] >> 64, st["state"]["state"] & m, st["state"]["inc"] >> 64, st["state"]["inc"] & m,
                st["has_uint32"], st["uinteger"]]
        assert [int(x) for x in rng[i]] == want


def test_mixed_draw_streams_match_numpy():
    nseeds = 200
    o = OracleVecEnv("MiniGrid-Empty-5x5-v0", nseeds)
    o.seed(np.arange(nseeds, dtype=np.uint64) * 7919)
    plan = np.random.default_rng(5)
    for i in range(nseeds):
        g = _gen(i * 7919)
        for _ in range(40):
            op = plan.integers(0, 3)
            if op == 0:
                lo = int(plan.integers(-5, 5)); hi = lo + int(plan.integers(1, 40))
                assert o.rng_integers(i, lo, hi) == int(g.integers(lo, hi))
            elif op == 1:
                n = int(plan.integers(1, 12))
                lst = list(range(n))
                g.shuffle(lst)
                assert list(o.rng_shuffle(i, n)) == lst
            else:
                a = int(plan.integers(0, 9)); b = a + int(plan.integers(1, 9))
                assert a + o.rng_integers(i, 0, b - a) == int(g.choice(range(a, b)))
    # the buffered 32-bit halves
    o2 = OracleVecEnv("MiniGrid-Empty-5x5-v0", 1)
    o2.seed(np.array([42], dtype=np.uint64))
    bg = np.random.PCG64(np.random.SeedSequence(42))
    raw = bg.random_raw(3)
    got = [o2.rng_next32(0) for _ in range(6)]
    want = []
    for r in raw:
        want += [int(r) & 0xFFFFFFFF, int(r) >> 32]
    assert got == want
