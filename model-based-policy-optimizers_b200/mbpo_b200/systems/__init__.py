"""mbpo.systems work-alike (mbpo/systems/__init__.py:1-4)."""
from .base_systems import System, SystemParams, SystemState
from .brax_wrapper import BraxWrapper
from .general_systems import NoisyPendulumSystem, PointMassParams, PointMassSystem
from .mlp_ensemble_system import MLPEnsembleSystem, MlpEnsembleDynamicsParams
from .pendulum_system import (PendulumDynamics, PendulumDynamicsParams, PendulumReward, PendulumRewardParams,
                              PendulumSystem)

__all__ = ["BraxWrapper", "NoisyPendulumSystem", "PointMassParams", "PointMassSystem", "MLPEnsembleSystem", "MlpEnsembleDynamicsParams", "System", "SystemParams", "SystemState", "PendulumSystem", "PendulumDynamics", "PendulumDynamicsParams",
           "PendulumReward", "PendulumRewardParams"]
