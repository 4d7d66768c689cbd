from .icem_optimizer import AbstractCost, iCEMOptimizer, iCemOptimizerState, iCemParams, iCemTO, iCemTrainingOutput

__all__ = ["AbstractCost", "iCEMOptimizer", "iCemOptimizerState", "iCemParams", "iCemTO", "iCemTrainingOutput"]
