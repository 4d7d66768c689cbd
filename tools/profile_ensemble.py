"""Launches the fused ensemble rollout kernel a few times at the config-4 shape (for ncu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import bench  # noqa: E402
from mbpo_b200.systems import MLPEnsembleSystem, MlpEnsembleDynamicsParams, PendulumRewardParams, SystemParams  # noqa: E402

dev = torch.device("cuda", 0)
ws, bs = bench.make_ensemble_numpy()
dyn = MlpEnsembleDynamicsParams(weights=[torch.from_numpy(w).to(dev) for w in ws],
                                biases=[torch.from_numpy(b).to(dev) for b in bs])
system = MLPEnsembleSystem()
sp = SystemParams(dynamics_params=dyn, reward_params=PendulumRewardParams())
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.ENS_B
x0 = torch.from_numpy(bench.random_states(B, 0)).to(dev)
acts = torch.zeros((B, bench.ENS_N + 15, bench.ENS_H, 1), device=dev).uniform_(-1, 1)
for _ in range(3):
    out = system.ensemble_returns(sp, x0, acts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    out = system.ensemble_returns(sp, x0, acts)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flops = B * (bench.ENS_N + 15) * bench.ENS_E * bench.ENS_H * bench.ENS_FLOP_PER_FORWARD
print("ensemble_rollout_kernel: %.3f ms, %.1f TFLOP/s, mean return %.4f" % (ms, flops / ms / 1e9, float(out.mean())))
