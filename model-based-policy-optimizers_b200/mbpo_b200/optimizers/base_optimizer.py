"""BaseOptimizer API shell (mbpo/optimizers/base_optimizer.py:14-57)."""
from __future__ import annotations

from typing import Tuple

import torch

from ..systems.base_systems import System
from ..utils.type_aliases import OptimizerState, OptimizerTrainingOutPut


class BaseOptimizer:
    def __init__(self, system: System | None = None, key: torch.Tensor | None = None):
        self.system = system
        self.key = key

    def set_system(self, system: System):
        self.system = system

    @property
    def can_act_in_batches(self):
        return True

    def act(self, obs: torch.Tensor, opt_state: OptimizerState, evaluate: bool = True) -> Tuple[torch.Tensor, OptimizerState]:
        raise NotImplementedError

    def train(self, opt_state: OptimizerState) -> OptimizerTrainingOutPut:
        return OptimizerTrainingOutPut(optimizer_state=opt_state)

    def init(self, key: torch.Tensor, true_buffer_state=None) -> OptimizerState:
        raise NotImplementedError

    def dummy_true_buffer_state(self, key: torch.Tensor):
        """base_optimizer.py:44-57: a 10-slot brax UniformSamplingQueue of dummy Transitions
        (observation [x_dim], action [u_dim], reward [1], discount [1], next_observation [x_dim]), ``.init(key)``.
        The state is what ``BraxWrapper(sample_buffer_state=...)`` expects.  A batch of keys [B, 2] (the additive
        vmapped ``init``) gives the vmapped pytree: ring [B, 10, D], key [B, 2]."""
        assert self.system is not None, "Base optimizer requires system to be defined."
        from ..replay_buffers import UniformSamplingQueue
        from ..utils.optimizer_utils import Transition
        dev = key.device
        z = lambda n: torch.zeros((n,), dtype=torch.float32, device=dev)
        dummy_transition = Transition(observation=z(self.system.x_dim), action=z(self.system.u_dim), reward=z(1),
                                      discount=z(1), next_observation=z(self.system.x_dim))
        sampling_buffer = UniformSamplingQueue(max_replay_size=10, dummy_data_sample=dummy_transition,
                                               sample_batch_size=1)
        if key.dim() == 1:
            return sampling_buffer.init(key)
        batch = tuple(key.shape[:-1])
        flat = key.reshape(-1, 2)
        state = sampling_buffer.init(flat[0] if flat.shape[0] else torch.zeros((2,), dtype=key.dtype, device=dev))
        return state.replace(ring=state.ring.expand(batch + tuple(state.ring.shape)).contiguous(), key=key)
