import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import bench, mbpo_b200
from mbpo_b200.envs import wrap
from mbpo_b200.systems import PendulumSystem
dev = torch.device("cuda", 0)
E, T = 65536, 1000
system = PendulumSystem()
env = wrap(system, system.reset(device=dev).system_params, episode_length=200)
st = env.reset(torch.from_numpy(bench.random_states(E, 1)).to(dev))
acts_host = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (T, E, 1)).astype(np.float32)).pin_memory()
rew_host = torch.empty((T, E), dtype=torch.float32).pin_memory()
def timeit(f, reps=6):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        out = f(); torch.cuda.synchronize(); del out
    return (time.perf_counter() - t0) / reps * 1e3
for chunk in (1000, 250, 125, 50):
    print("chunk %4d: H2D+rollout %.2f ms, +D2H rewards %.2f ms" % (
        chunk, timeit(lambda: env.unroll_streamed(st, acts_host, None, chunk)),
        timeit(lambda: env.unroll_streamed(st, acts_host, rew_host, chunk))), flush=True)
a_dev = acts_host.to(dev)
print("rollout alone (device actions): %.2f ms" % timeit(lambda: env.unroll(st, a_dev)))
print("chunked rollouts alone:", ["%d: %.2f ms" % (c, timeit(lambda: [env.unroll(st, a_dev[t0:t0 + c]) for t0 in range(0, T, c)])) for c in (250, 125)])
