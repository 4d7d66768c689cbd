"""Two launches of the wide (latency) actor kernel at a given env count.  With the library built by
`MBPO_EXTRA_NVCC_FLAGS=-DMBPO_ATCW_PROFILE python model-based-policy-optimizers_b200/build.py --force` the kernel prints
its per-step phase clocks (clock64); it is also the target of the ncu capture in profiles/."""
import os, subprocess, sys
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
policy = acting.Policy(acting.PolicyParams([torch.from_numpy(w).to(dev) for w in pol_w],
                                           [torch.from_numpy(b).to(dev) for b in pol_b]), kernel="tcgen05_wide")
key = mbpo_b200.random.PRNGKey(0, dev)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
st = env.reset(torch.from_numpy(bench.random_states(E, 1)).to(dev))
for _ in range(2):
    acting.get_experience(env, st, policy, key, 200)
torch.cuda.synchronize()
