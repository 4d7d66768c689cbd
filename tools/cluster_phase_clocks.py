"""Phase clocks (clock64 of thread 0 of CTA 0, second iCEM iteration) of the cluster plan at B = 1.
Needs a library built with -DMBPO_CLUSTER_CLOCKS:  tools/ab_variant.sh clk 30 -DMBPO_CLUSTER_CLOCKS  and
MBPO_B200_LIB=<that .so> python tools/cluster_phase_clocks.py [cluster]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import numpy as np, torch
import mbpo_b200
from mbpo_b200 import _lib
from mbpo_b200.optimizers import iCemTO, iCemParams
from mbpo_b200.systems import PendulumSystem
dev = torch.device("cuda", 0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
opt = iCemTO(horizon=30, action_dim=1, opt_params=iCemParams(num_samples=512, num_particles=1)); opt.set_system(PendulumSystem())
st = opt.init(mbpo_b200.random.split(mbpo_b200.random.PRNGKey(0, dev), 1))
x0 = torch.tensor([[-1.0, 0.0, 0.0]], device=dev)
for _ in range(5): opt._plan_raw(x0, st.key, st.best_sequence, st.system_params, cluster=C)
torch.cuda.synchronize()
lib = _lib.lib
buf = (ctypes.c_longlong * 64)()
lib.mbpo_debug_cluster_clocks.argtypes = [ctypes.c_void_p]; lib.mbpo_debug_cluster_clocks(buf)
c = list(buf)[:10]; x = list(buf)[10:13]
if x[2] > x[0] > c[5]: print("ranked selection detail: entry", x[0] - c[5], "compare", x[1] - x[0], "finalize", x[2] - x[1], "barrier", c[6] - x[2])
names = ["0 -", "1 -", "2 rollout", "3 push keys", "4 cluster.sync", "5 selection",
         "6 push elites", "7 cluster.sync", "8 refit + apply"]
print(json.dumps({"cluster": C, "cycles": {names[i]: c[i + 1] - c[i] for i in range(9)}, "iteration": c[9] - c[0]}))
sbuf = (ctypes.c_longlong * 16)()
lib.mbpo_debug_select_clocks.argtypes = [ctypes.c_void_p]; lib.mbpo_debug_select_clocks(sbuf)
sc = list(sbuf)[:7]
if sc[6] > sc[0] > 0:
    print(json.dumps({"cta_select phases (last call)": {n: sc[i + 1] - sc[i] for i, n in enumerate(
        ["0 reset + or/and", "1 histogram", "2 boundary bin", "3 classify", "4 ranks", "5 -"])}}))
