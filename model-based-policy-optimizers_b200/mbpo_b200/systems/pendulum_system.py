"""Analytic pendulum swing-up System.

Mirrors mbpo/systems/pendulum_system.py:12-46, dynamics/pendulum_dynamics.py:12-63 and
rewards/pendulum_reward.py:12-42.  State [cos th, sin th, thdot], action 1-D.  ``step`` runs
the CUDA kernel behind ``mbpo_system_step``.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .. import _lib
from ..config import config
from .base_systems import System, SystemParams, SystemState, _Replaceable


@dataclass
class PendulumDynamicsParams(_Replaceable):
    """pendulum_dynamics.py:12-19."""
    max_speed: float = 8.0
    max_torque: float = 2.0
    dt: float = 0.05
    g: float = 9.81
    m: float = 1.0
    l: float = 1.0


@dataclass
class PendulumRewardParams(_Replaceable):
    """pendulum_reward.py:12-16."""
    control_cost: float = 0.02
    angle_cost: float = 1.0
    target_angle: float = 0.0


class PendulumDynamics:
    x_dim, u_dim = 3, 1

    def init_params(self, key) -> PendulumDynamicsParams:
        return PendulumDynamicsParams()


class PendulumReward:
    x_dim, u_dim = 3, 1

    def init_params(self, key) -> PendulumRewardParams:
        return PendulumRewardParams()


def pack_pendulum(dynamics_params, reward_params) -> _lib.PendulumParamsC:
    d = dynamics_params if dynamics_params is not None else PendulumDynamicsParams()
    r = reward_params if reward_params is not None else PendulumRewardParams()
    return _lib.PendulumParamsC(float(d.max_speed), float(d.max_torque), float(d.dt), float(d.g), float(d.m),
                                float(d.l), float(r.control_cost), float(r.angle_cost), float(r.target_angle))


class PendulumSystem(System):
    system_kind = _lib.SYSTEM_PENDULUM

    def __init__(self):
        super().__init__(dynamics=PendulumDynamics(), reward=PendulumReward())
        self.min_action = -1.0
        self.max_action = 1.0

    def pack_params(self, system_params: SystemParams):
        return pack_pendulum(system_params.dynamics_params, system_params.reward_params)

    def step(self, x: torch.Tensor, u: torch.Tensor, system_params: SystemParams) -> SystemState:
        """x[..., 3], u[..., 1] -> SystemState(x_next[..., 3], reward[...]).  Like the reference
        (pendulum_system.py:38) the returned SystemParams carries no key."""
        if x.shape[-1] != 3 or u.shape[-1] != 1 or x.shape[:-1] != u.shape[:-1]:
            raise ValueError("PendulumSystem.step expects x[..., 3] and u[..., 1] with equal batch dims")
        xc = x.to(torch.float32).contiguous()
        uc = u.to(torch.float32).contiguous()
        rows = xc.numel() // 3
        x_next = torch.empty_like(xc)
        reward = torch.empty(xc.shape[:-1], dtype=torch.float32, device=xc.device)
        params = self.pack_params(system_params)
        with _lib.cuda_guard(xc):
            _lib.check(_lib.lib.mbpo_system_step(self.system_kind, _lib.C.addressof(params), config.math_mode_id,
                                                 _lib.ptr(xc), _lib.ptr(uc), rows, _lib.ptr(x_next),
                                                 _lib.ptr(reward), _lib.stream_ptr(xc.device)))
        return SystemState(x_next=x_next, reward=reward,
                           system_params=SystemParams(dynamics_params=system_params.dynamics_params,
                                                      reward_params=system_params.reward_params))

    def reset(self, rng: torch.Tensor = None, device=None) -> SystemState:
        """pendulum_system.py:41-46: x = [-1, 0, 0] (hanging down), reward 0.  A batch of
        keys [..., 2] gives a batch of states (= jax.vmap(system.reset))."""
        dev = rng.device if rng is not None else _lib.require_cuda(device)
        batch = tuple(rng.shape[:-1]) if rng is not None else ()
        x = torch.tensor([-1.0, 0.0, 0.0], dtype=torch.float32, device=dev).expand(batch + (3,)).contiguous()
        return SystemState(x_next=x, reward=torch.zeros(batch, dtype=torch.float32, device=dev),
                           system_params=SystemParams(dynamics_params=PendulumDynamicsParams(),
                                                      reward_params=PendulumRewardParams()))

