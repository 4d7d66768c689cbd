"""Wall time per get_experience call in the reference's own regime (tests/test_sac.py: 32 envs, 20 steps per call):
eager host path against the CUDA-graph replay."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import bench, mbpo_b200
from mbpo_b200 import acting
from mbpo_b200.envs import wrap
from mbpo_b200.systems import PendulumSystem
dev = torch.device("cuda", 0)
pol_w, pol_b = bench.make_policy_numpy(seed=7)
system = PendulumSystem()
env = wrap(system, system.reset(device=dev).system_params, episode_length=200)
policy = acting.Policy(acting.PolicyParams([torch.from_numpy(w).to(dev) for w in pol_w],
                                           [torch.from_numpy(b).to(dev) for b in pol_b]))
key = mbpo_b200.random.PRNGKey(0, dev)
for E, T in ((32, 20), (128, 20), (1024, 20), (32, 200)):
    st = env.reset(torch.from_numpy(bench.random_states(E, 1)).to(dev))
    def eager(n, k=key, s=st):
        for _ in range(n):
            k, s, tr = acting.get_experience(env, s, policy, k, T)
        torch.cuda.synchronize()
    collect = acting.GraphedRollout(env, st, policy, key, T)
    def graphed(n):
        for _ in range(n):
            collect()
        torch.cuda.synchronize()
    res = {}
    for name, f in (("eager", eager), ("graph", graphed)):
        f(20)
        t0 = time.perf_counter(); f(200); res[name] = (time.perf_counter() - t0) / 200 * 1e6
    print(json.dumps({"envs": E, "steps_per_call": T, "eager_us_per_call": res["eager"], "graph_us_per_call": res["graph"],
                      "graph_us_per_policy_step": res["graph"] / T}), flush=True)
