from .optimizer_utils import Transition, rollout_actions, rollout_returns
from .type_aliases import OptimizerState, OptimizerTrainingOutPut

__all__ = ["Transition", "rollout_actions", "rollout_returns", "OptimizerState", "OptimizerTrainingOutPut"]
