from .optimizer_utils import (Transition, compute_gae, lambda_return, lambda_return_vjp, rollout_actions, rollout_policy,
                              rollout_policy_vjp, rollout_returns)
from .type_aliases import OptimizerState, OptimizerTrainingOutPut

__all__ = ["Transition", "rollout_actions", "rollout_returns", "rollout_policy", "rollout_policy_vjp", "lambda_return",
           "lambda_return_vjp", "compute_gae", "OptimizerState", "OptimizerTrainingOutPut"]
