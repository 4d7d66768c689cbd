"""How many clusters of a size run at once: plan latency against the number of problems for each cluster size."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import numpy as np, torch
import mbpo_b200
from mbpo_b200.optimizers import iCemTO, iCemParams
from mbpo_b200.systems import PendulumSystem
dev = torch.device("cuda", 0)
def timeit(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
opt = iCemTO(horizon=30, action_dim=1, opt_params=iCemParams(num_samples=512, num_particles=1)); opt.set_system(PendulumSystem())
for c, Bs in ((16, (1, 2, 4, 7, 8, 9, 10)), (8, (8, 14, 15, 16, 17, 18, 19)), (4, (18, 24, 30, 32, 34, 36, 37, 38)), (2, (37, 56, 64, 70, 72, 73, 74, 75))):
    row = {"cluster": c}
    for B in Bs:
        st = opt.init(mbpo_b200.random.split(mbpo_b200.random.PRNGKey(0, dev), B))
        rng = np.random.default_rng(0); th, w = rng.uniform(-np.pi, np.pi, B), rng.uniform(-8, 8, B)
        x0 = torch.from_numpy(np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)).to(dev)
        row["B=%d" % B] = round(timeit(lambda: opt._plan_raw(x0, st.key, st.best_sequence, st.system_params, cluster=c)) * 1e3, 1)
    print(json.dumps(row), flush=True)
