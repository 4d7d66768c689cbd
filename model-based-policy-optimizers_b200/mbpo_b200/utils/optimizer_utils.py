"""rollout_actions: open-loop rollout of action sequences through System.step.

Mirrors mbpo/utils/optimizer_utils.py:11-59 and returns the same brax-style Transition
fields.  Batched forms replace ``jax.vmap(rollout_actions)``:
    init_state [X],   actions [H, A]        -> fields [H, ...]
    init_state [B,X], actions [B, M, H, A]  -> fields [B, M, H, ...]  (M sequences per state)
"""
from __future__ import annotations

from typing import Any, NamedTuple

import torch

from .. import _lib
from ..config import config
from ..systems.base_systems import System, SystemParams


class Transition(NamedTuple):
    """brax.training.types.Transition."""
    observation: torch.Tensor
    action: torch.Tensor
    reward: torch.Tensor
    discount: torch.Tensor
    next_observation: torch.Tensor
    extras: Any = ()


def rollout_actions(system: System, system_params: SystemParams, init_state: torch.Tensor, actions: torch.Tensor,
                    horizon: int) -> Transition:
    single = init_state.dim() == 1
    if single:
        if actions.dim() != 2:
            raise ValueError("rollout_actions: actions must be [H, A] for a single init_state")
        x0 = init_state.reshape(1, -1)
        acts = actions.reshape(1, 1, *actions.shape)
    else:
        if actions.dim() != 4 or actions.shape[0] != init_state.shape[0]:
            raise ValueError("rollout_actions: batched form needs init_state [B, X] and actions [B, M, H, A]")
        x0, acts = init_state, actions
    assert acts.shape[2] == horizon, "actions.shape[0] must equal horizon"   # optimizer_utils.py:26
    x0 = x0.to(torch.float32).contiguous()
    acts = acts.to(torch.float32).contiguous()
    B, M, H, A = acts.shape
    X = x0.shape[-1]
    dev = x0.device
    obs = torch.empty((B, M, H, X), dtype=torch.float32, device=dev)
    nxt = torch.empty((B, M, H, X), dtype=torch.float32, device=dev)
    rew = torch.empty((B, M, H), dtype=torch.float32, device=dev)
    params = system.pack_params(system_params)
    with _lib.cuda_guard(x0):
        _lib.check(_lib.lib.mbpo_rollout_actions(system.system_kind, _lib.C.addressof(params), config.math_mode_id, H, A,
                                                 X, _lib.ptr(x0), _lib.ptr(acts), B, M, None, _lib.ptr(obs),
                                                 _lib.ptr(rew), _lib.ptr(nxt), _lib.stream_ptr(dev)))
    tr = Transition(observation=obs, action=acts, reward=rew, discount=torch.ones_like(rew), next_observation=nxt)
    if single:
        tr = Transition(*(f[0, 0] for f in tr[:5]))
    return tr


def rollout_returns(system: System, system_params: SystemParams, init_state: torch.Tensor,
                    actions: torch.Tensor) -> torch.Tensor:
    """Horizon-mean reward of each action sequence (icem_optimizer.py:160 inner mean) without
    materialising the Transition: init_state [B,X], actions [B,M,H,A] -> [B,M]."""
    x0 = init_state.to(torch.float32).contiguous()
    acts = actions.to(torch.float32).contiguous()
    B, M, H, A = acts.shape
    out = torch.empty((B, M), dtype=torch.float32, device=x0.device)
    params = system.pack_params(system_params)
    with _lib.cuda_guard(x0):
        _lib.check(_lib.lib.mbpo_rollout_actions(system.system_kind, _lib.C.addressof(params), config.math_mode_id, H, A,
                                                 x0.shape[-1], _lib.ptr(x0), _lib.ptr(acts), B, M, _lib.ptr(out), None,
                                                 None, None, _lib.stream_ptr(x0.device)))
    return out
