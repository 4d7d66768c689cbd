"""mbpo_b200: B200-native iCEM planning hot path behind the mbpo Python API.

    from mbpo_b200.optimizers import iCemTO, iCemParams
    from mbpo_b200.systems import PendulumSystem
    import mbpo_b200.random as jr

Importing the package loads libmbpo_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/mbpo_b200.h); it raises if the library is not built.  There is no CPU fallback.
"""
from . import _lib, acting, envs, optimizers, parallel, random, replay_buffers, running_statistics, systems, utils
from .config import config
from ._lib import MbpoError, MbpoUnsupported

__all__ = ["config", "acting", "envs", "optimizers", "parallel", "random", "replay_buffers", "running_statistics", "systems", "utils", "MbpoError", "MbpoUnsupported"]
__version__ = "0.1.0"
