"""mbpo.optimizers work-alike (mbpo/optimizers/__init__.py:1-6), planning path only."""
from .base_optimizer import BaseOptimizer
from .trajectory_optimizers.icem_optimizer import (AbstractCost, iCEMOptimizer, iCemOptimizerState, iCemParams,
                                                   iCemTO, iCemTrainingOutput)

__all__ = ["BaseOptimizer", "AbstractCost", "iCEMOptimizer", "iCemOptimizerState", "iCemParams", "iCemTO",
           "iCemTrainingOutput"]
