"""System interface: the drop-in boundary of the planning hot path.

Mirrors mbpo/systems/base_systems.py:13-60 (SystemParams, SystemState, System) with torch
CUDA tensors in place of jax arrays.  Where the reference writes ``jax.vmap(system.step)``,
call ``system.step`` with leading batch dimensions: the CUDA kernels are the vmapped form.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Any, Optional

import torch

from .. import _lib
from .. import random as jr


class _Replaceable:
    """chex.dataclass-style functional update (used at icem_optimizer.py:147,247,251)."""

    def replace(self, **changes):
        return dataclasses.replace(self, **changes)


@dataclass
class SystemParams(_Replaceable):
    """base_systems.py:13-17.  ``key`` defaults to PRNGKey(0), created lazily on first use."""
    dynamics_params: Any = None
    reward_params: Any = None
    key: Optional[torch.Tensor] = None


@dataclass
class SystemState(_Replaceable):
    """base_systems.py:20-25."""
    x_next: torch.Tensor = None
    reward: torch.Tensor = None
    system_params: SystemParams = None
    done: Any = 0.0


class System:
    """base_systems.py:28-60.  Subclasses advertise ``system_kind`` (which inlined CUDA
    System.step the kernels use) and pack their parameters for the C ABI."""

    system_kind: int = -1

    def __init__(self, dynamics=None, reward=None, x_dim: int = 0, u_dim: int = 0):
        self.dynamics = dynamics
        self.reward = reward
        self.x_dim = dynamics.x_dim if dynamics is not None else x_dim
        self.u_dim = dynamics.u_dim if dynamics is not None else u_dim

    @staticmethod
    def system_params_vmap_axes(axes: int = 0):
        return SystemParams(dynamics_params=None, reward_params=None, key=axes)

    def step(self, x: torch.Tensor, u: torch.Tensor, system_params: SystemParams) -> SystemState:
        raise NotImplementedError

    def init_params(self, key: torch.Tensor) -> SystemParams:
        keys = jr.split(key, 3)
        return SystemParams(dynamics_params=self.dynamics.init_params(keys[..., 0, :]),
                            reward_params=self.reward.init_params(keys[..., 1, :]),
                            key=keys[..., 2, :].contiguous())

    # ---- C-ABI side -------------------------------------------------------------------------
    def pack_params(self, system_params: SystemParams):
        """ctypes struct (kept alive by the caller) passed as ``sys_params_host``."""
        raise _lib.MbpoUnsupported(_lib.MBPO_EUNSUPPORTED,
                                   "%s has no inlined CUDA System.step (no fallback path exists)" % type(self).__name__)
