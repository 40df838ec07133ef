"""minigrid_b200 — a B200-native, lockstep-batched engine for the Minigrid hot path
(MiniGridEnv.step / reset / gen_obs for a batch of independent environments).

    from minigrid_b200 import MinigridVecEnv
    envs = MinigridVecEnv("MiniGrid-DoorKey-8x8-v0", num_envs=262144)
    obs, _ = envs.reset(seed=0)
    obs, reward, terminated, truncated, _ = envs.step(actions)   # torch CUDA tensors

The compute path is hand-written CUDA for sm_100a behind the C-ABI in include/minigrid_b200.h; there is no
CPU implementation in this package.
"""
from . import specs  # noqa: F401
from ._lib import MinigridB200Error  # noqa: F401
from ._numa import bind_to_gpu_numa_node, gpu_numa_node  # noqa: F401
from .vector_env import MinigridVecEnv, make_sharded, shard_range  # noqa: F401
from .wrappers import (ActionBonus, DictObservationSpaceWrapper, FlatObsWrapper, FullyObsWrapper, ImgObsWrapper, NoDeath, OneHotPartialObsWrapper,  # noqa: F401
                       PositionBonus, RGBImgObsWrapper, RGBImgPartialObsWrapper, SymbolicObsWrapper, ViewSizeWrapper)

__version__ = "0.1.0"
