"""cta_select phase clocks (clock64 of thread 0 of CTA 0, last call) inside the fused plan at config 2's size, i.e. with
three CTAs per SM competing.  Needs a library built with -DMBPO_CLUSTER_CLOCKS (tools/ab_variant.sh clk 30 ...)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import numpy as np, torch
import mbpo_b200
from mbpo_b200 import _lib
from mbpo_b200.optimizers import iCemTO, iCemParams
from mbpo_b200.systems import PendulumSystem
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
opt = iCemTO(horizon=30, action_dim=1, opt_params=iCemParams(num_samples=512, num_particles=1)); opt.set_system(PendulumSystem())
st = opt.init(mbpo_b200.random.split(mbpo_b200.random.PRNGKey(0, dev), B))
rng = np.random.default_rng(0); th, w = rng.uniform(-np.pi, np.pi, B), rng.uniform(-8, 8, B)
x0 = torch.from_numpy(np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)).to(dev)
for _ in range(3): opt._plan_raw(x0, st.key, st.best_sequence, st.system_params, cluster=1)
torch.cuda.synchronize()
sbuf = (ctypes.c_longlong * 16)()
_lib.lib.mbpo_debug_select_clocks.argtypes = [ctypes.c_void_p]; _lib.lib.mbpo_debug_select_clocks(sbuf)
sc = list(sbuf)[:7]
print(json.dumps({"problems": B, "cta_select phases": {n: sc[i + 1] - sc[i] for i, n in enumerate(
    ["0 reset + or/and", "1 histogram", "2 boundary bin", "3 classify", "4 ranks", "5 -"])}, "total": sc[6] - sc[0]}))
