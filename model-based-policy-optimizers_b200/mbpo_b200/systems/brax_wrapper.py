"""BraxWrapper (mbpo/systems/brax_wrapper.py:14-66): a System as a brax env whose resets draw the first observation
from the true replay buffer.  ``reset`` accepts one key [2] or a batch of keys [E, 2] (= VmapWrapper.reset,
brax_utils/training.py:66-69) and runs one launch of ``mbpo_env_reset_from_buffer``."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional

import torch

from .. import _lib
from ..config import config
from ..replay_buffers import ReplayBufferState, UniformSamplingQueue
from .base_systems import System, SystemParams, _Replaceable


@dataclass
class State(_Replaceable):
    """brax_utils/base.py:12-23."""
    pipeline_state: Any = None
    obs: torch.Tensor = None
    reward: torch.Tensor = None
    done: torch.Tensor = None
    system_params: Optional[SystemParams] = None


class BraxWrapper:
    def __init__(self, system: System, system_params: SystemParams, sample_buffer_state: ReplayBufferState,
                 sample_buffer: UniformSamplingQueue):
        self.system = system
        self.sample_buffer_state = sample_buffer_state
        self.sample_buffer = sample_buffer
        self.init_system_params = system_params

    def reset(self, rng: torch.Tensor) -> State:
        single = rng.dim() == 1
        rngs = rng.reshape(-1, 2).contiguous()
        E, X = rngs.shape[0], self.system.x_dim
        dev = rngs.device
        obs = torch.empty((E, X), dtype=torch.float32, device=dev)
        reward = torch.empty((E,), dtype=torch.float32, device=dev)
        keys = torch.empty((E, 2), dtype=torch.uint32, device=dev)
        st = self.sample_buffer_state._c()
        # a row is ravel_pytree(Transition): observation [X], action [A], reward, ...
        reward_col = self.sample_buffer.column_of(2)
        with _lib.cuda_guard(rngs):
            _lib.check(_lib.lib.mbpo_env_reset_from_buffer(
                _lib.C.byref(st), _lib.ptr(rngs), E, config.prng_mode, self.sample_buffer._sample_batch_size, X,
                reward_col, _lib.ptr(obs), _lib.ptr(reward), _lib.ptr(keys), None, _lib.stream_ptr(dev)))
        done = torch.zeros((E,), dtype=torch.float32, device=dev)
        if single:
            obs, reward, done, keys = obs[0], reward[0], done[0], keys[0]
        return State(pipeline_state=None, obs=obs, reward=reward, done=done,
                     system_params=self.init_system_params.replace(key=keys))

    def step(self, state: State, action: torch.Tensor) -> State:
        nxt = self.system.step(state.obs, action, state.system_params)
        done = nxt.done if torch.is_tensor(nxt.done) else torch.full_like(nxt.reward, float(nxt.done))
        return state.replace(obs=nxt.x_next, reward=nxt.reward, done=done, system_params=nxt.system_params)

    @property
    def action_size(self) -> int:
        return self.system.u_dim

    @property
    def observation_size(self) -> int:
        return self.system.x_dim

    @property
    def backend(self) -> str:
        return "string"
