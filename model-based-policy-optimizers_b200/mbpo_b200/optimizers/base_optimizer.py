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
        """The reference builds a 10-slot brax UniformSamplingQueue here (base_optimizer.py:44-57);
        the planning path only carries it.  We carry the key so the field stays non-empty."""
        assert self.system is not None, "Base optimizer requires system to be defined."
        return {"key": key}
