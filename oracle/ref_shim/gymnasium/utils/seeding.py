"""seeding.np_random stand-in: Generator(PCG64(SeedSequence(seed)))."""
from __future__ import annotations

import numpy as np

RandomNumberGenerator = np.random.Generator
RNG = RandomNumberGenerator


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError(f"Seed must be a non-negative integer, got {seed!r}")
    seed_seq = np.random.SeedSequence(seed)
    return np.random.Generator(np.random.PCG64(seed_seq)), seed_seq.entropy
