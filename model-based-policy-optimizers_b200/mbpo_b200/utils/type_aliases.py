"""mbpo/utils/type_aliases.py:10-19 work-alike."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import torch

from ..systems.base_systems import SystemParams, _Replaceable


@dataclass
class OptimizerState(_Replaceable):
    true_buffer_state: Any = None          # opaque (brax ReplayBufferState in the reference)
    system_params: SystemParams = None
    key: torch.Tensor = None


@dataclass
class OptimizerTrainingOutPut(_Replaceable):
    optimizer_state: OptimizerState = None
