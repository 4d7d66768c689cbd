"""mbpo.systems work-alike (mbpo/systems/__init__.py:1-4)."""
from .base_systems import System, SystemParams, SystemState
from .pendulum_system import (PendulumDynamics, PendulumDynamicsParams, PendulumReward, PendulumRewardParams,
                              PendulumSystem)

__all__ = ["System", "SystemParams", "SystemState", "PendulumSystem", "PendulumDynamics", "PendulumDynamicsParams",
           "PendulumReward", "PendulumRewardParams"]
