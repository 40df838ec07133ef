"""Minimal stand-in for the `gymnasium` package (TEST INFRASTRUCTURE ONLY).

gymnasium is not installed in this image and cannot be fetched (no network). The
unmodified reference at /root/reference imports it at module scope
(minigrid/minigrid_env.py:8-13, minigrid/__init__.py:3), so this shim supplies
just the names the reference touches on the step/reset/gen_obs path. It is our
own code (nothing is copied from gymnasium); semantics that matter for parity:

* ``Env.reset(seed=s)`` re-creates ``np_random = Generator(PCG64(SeedSequence(s)))``
  exactly like ``gymnasium.utils.seeding.np_random``; ``reset()`` keeps the stream.
* A real gymnasium install, if present earlier on sys.path, always wins.
"""
from __future__ import annotations

from . import logger, spaces  # noqa: F401
from .core import ActionWrapper, Env, ObservationWrapper, RewardWrapper, Wrapper  # noqa: F401
from .envs.registration import make, register, registry, spec  # noqa: F401

__version__ = "0.0-shim"
