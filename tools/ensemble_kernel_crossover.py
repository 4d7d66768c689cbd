"""One-tile vs two-tile (ping-pong) ensemble rollout kernel around the row count where the library switches
(2 x 128 x 74 = 18,944 rows): the same work either side of the threshold, timed with CUDA events."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import numpy as np, torch
import mbpo_b200
from mbpo_b200.systems import MLPEnsembleSystem, MlpEnsembleDynamicsParams, PendulumRewardParams, SystemParams
dev = torch.device("cuda", 0)
rng = np.random.default_rng(3)
dims = (4, 256, 256, 256, 3)
ws = [torch.from_numpy((rng.standard_normal((5, dims[i], dims[i + 1])) / np.sqrt(dims[i])).astype(np.float32)).to(dev) for i in range(4)]
bs = [torch.from_numpy((0.01 * rng.standard_normal((5, dims[i + 1]))).astype(np.float32)).to(dev) for i in range(4)]
system = MLPEnsembleSystem()
sp = SystemParams(dynamics_params=MlpEnsembleDynamicsParams(weights=ws, biases=bs), reward_params=PendulumRewardParams())
H = 50
for B, M in ((1, 1039), (4, 1039), (9, 1039), (18, 1052), (18, 1053), (36, 1039)):
    x0 = torch.zeros((B, 3), device=dev); x0[:, 0] = -1
    acts = torch.zeros((B, M, H, 1), device=dev).uniform_(-1, 1)
    for _ in range(3): system.ensemble_returns(sp, x0, acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): system.ensemble_returns(sp, x0, acts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    rows = B * M
    print(json.dumps({"rows": rows, "kernel": "one tile per CTA pair" if rows <= 18944 else "two tiles (ping-pong)",
                      "ms": round(ms, 4), "TFLOPs": round(rows * 5 * H * 265728 / ms / 1e9, 1)}), flush=True)
