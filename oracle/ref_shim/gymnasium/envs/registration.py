"""register()/make() stand-ins: a dict of id -> (entry_point, kwargs); no wrappers added."""
from __future__ import annotations

import importlib
from dataclasses import dataclass, field


@dataclass
class EnvSpec:
    id: str
    entry_point: object
    kwargs: dict = field(default_factory=dict)
    max_episode_steps: int | None = None


registry: dict[str, EnvSpec] = {}


def register(id, entry_point=None, kwargs=None, max_episode_steps=None, **_ignored):
    registry[id] = EnvSpec(id, entry_point, dict(kwargs or {}), max_episode_steps)


def spec(id):
    return registry[id]


def make(id, **kwargs):
    s = registry[id]
    ep = s.entry_point
    if isinstance(ep, str):
        mod, attr = ep.split(":")
        ep = getattr(importlib.import_module(mod), attr)
    kw = dict(s.kwargs)
    kw.update(kwargs)
    env = ep(**kw)
    env.spec = s
    return env
