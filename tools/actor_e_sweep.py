"""Times the policy-in-the-loop rollout (acting.get_experience, T = 200) for several env counts on one GPU:
the per-step latency curve behind the strong-scaling numbers of config 3 with the policy in the loop."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import bench
import mbpo_b200
from mbpo_b200 import acting
from mbpo_b200.envs import wrap
from mbpo_b200.systems import PendulumSystem

dev = torch.device("cuda", 0)
pol_w, pol_b = bench.make_policy_numpy(seed=7)
system = PendulumSystem()
env = wrap(system, system.reset(device=dev).system_params, episode_length=200)
T = 200
for kernel in ("tcgen05_wide", "tcgen05", "cuda_cores"):
    policy = acting.Policy(acting.PolicyParams([torch.from_numpy(w).to(dev) for w in pol_w],
                                               [torch.from_numpy(b).to(dev) for b in pol_b]), kernel=kernel)
    key = mbpo_b200.random.PRNGKey(0, dev)
    for E in (32, 128, 1024, 8192, 16384, 18944, 32768, 65536):
        st = env.reset(torch.from_numpy(bench.random_states(E, 1)).to(dev))
        for _ in range(3):
            acting.get_experience(env, st, policy, key, T)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            acting.get_experience(env, st, policy, key, T)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(json.dumps({"kernel": kernel, "envs": E, "T": T, "ms_per_call": ms, "us_per_step": ms * 1e3 / T,
                          "env_steps_per_s": E * T / (ms * 1e-3)}), flush=True)
