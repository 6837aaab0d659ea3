"""
loco_mujoco_b200: B200-native batched `LocoEnv.step()`.

    from loco_mujoco_b200 import LocoEnv
    env = LocoEnv.make("UnitreeA1.simple", num_envs=4096)      # batched, torch.cuda tensors
    env = LocoEnv.make("UnitreeA1.simple")                     # drop-in single env, numpy float64
"""
__version__ = "0.1.0"

import os as _os

# One hardware work queue per CUDA stream (effective only if CUDA is not initialised yet): with the default of 8 connections
# two streams of a MixedBatch can alias to one queue, which serialises the members' kernels (mixed batches).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .environments import LocoEnv
from .environments.gymnasium import GymnasiumWrapper, VectorGymnasiumWrapper, make_gym


def get_all_task_names():
    return LocoEnv.get_all_task_names()
