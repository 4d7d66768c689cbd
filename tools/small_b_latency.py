"""Plan latency at small B (config 1 and the left end of config 5): one CTA per problem vs thread-block clusters."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import numpy as np, torch
import mbpo_b200
from mbpo_b200.optimizers import iCemTO, iCemParams
from mbpo_b200.systems import PendulumSystem
dev = torch.device("cuda", 0)
def states(n):
    rng = np.random.default_rng(0); th, w = rng.uniform(-np.pi, np.pi, n), rng.uniform(-8, 8, n)
    return torch.from_numpy(np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)).to(dev)
def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
opt = iCemTO(horizon=30, action_dim=1, opt_params=iCemParams(num_samples=512, num_particles=1)); opt.set_system(PendulumSystem())
for B in (1, 8, 64):
    st = opt.init(mbpo_b200.random.split(mbpo_b200.random.PRNGKey(0, dev), B)); x0 = states(B)
    row = {"workload": "config2 problem (N=512, H=30, S=5)", "problems": B}
    for c in (1, 2, 4, 8, 16, -1):
        try:
            ms = timeit(lambda: opt._plan_raw(x0, st.key, st.best_sequence, st.system_params, cluster=c), 50)
            row["cluster_%s_ms" % ("auto" if c < 0 else c)] = round(ms, 4)
            row["cluster_%s_Gtps" % ("auto" if c < 0 else c)] = round(B * 5 * 527 * 30 / ms / 1e6, 2)
        except Exception as e:
            row["cluster_%s_ms" % c] = str(e)[:60]
    print(json.dumps(row), flush=True)
# config 1: the reference's closed loop, 200 MPC steps
jr = mbpo_b200.random; ks = jr.split(jr.PRNGKey(0, dev), 3)
system = PendulumSystem(); s0 = system.reset(ks[2])
cem = iCemTO(horizon=20, action_dim=1, opt_params=iCemParams(), key=ks[0]); cem.set_system(system); st = cem.init(ks[1])
row = {"workload": "config1 closed loop, 200 steps (N=500, H=20, P=10)"}
for c in (1, 2, 4, 8, 16):
    ms = timeit(lambda: cem.closed_loop(s0.x_next, st, 200, cluster=c), 5)
    row["cluster_%d_ms" % c] = round(ms, 3); row["cluster_%d_ms_per_plan" % c] = round(ms / 200, 4)
row["sum_rewards"] = float(cem.closed_loop(s0.x_next, st, 200)[1].sum())
print(json.dumps(row), flush=True)
