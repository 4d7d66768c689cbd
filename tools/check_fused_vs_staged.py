"""A/B helper: fused plan vs the staged plan (same per-stage kernels) on the loaded library; prints equality and time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import numpy as np, torch
import mbpo_b200
from mbpo_b200.optimizers import iCemTO, iCemParams
from mbpo_b200.systems import PendulumSystem
L = mbpo_b200._lib
dev = torch.device("cuda", 0)
H, B = int(os.environ.get("H", "30")), 64
p = iCemParams(num_samples=512, num_particles=1)
opt = iCemTO(horizon=H, action_dim=1, opt_params=p); opt.set_system(PendulumSystem())
cfg = opt._cfg()
sp = PendulumSystem().reset(device=dev).system_params
rng = np.random.default_rng(0)
th, w = rng.uniform(-np.pi, np.pi, B), rng.uniform(-8, 8, B)
x0 = torch.from_numpy(np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)).to(dev)
keys = mbpo_b200.random.split(mbpo_b200.random.PRNGKey(0, dev), B)
seq = torch.zeros((B, H, 1), device=dev)
f_seq, f_val, f_key, _ = opt._plan_raw(x0, keys, seq, sp)
nbytes = L.lib.mbpo_icem_workspace_bytes(L.C.byref(cfg), B)
ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
s_seq, s_val, s_key = torch.empty_like(f_seq), torch.empty_like(f_val), torch.empty_like(f_key)
pp = PendulumSystem().pack_params(sp)
L.check(L.lib.mbpo_icem_plan_staged(L.C.byref(cfg), L.C.addressof(pp), L.ptr(x0), L.ptr(keys), L.ptr(seq), B,
                                    L.ptr(s_seq), L.ptr(s_val), L.ptr(s_key), L.ptr(ws), nbytes, L.stream_ptr(dev)))
torch.cuda.synchronize()
print(os.environ.get("MBPO_B200_LIB", "in-tree"), "H", H, "fused==staged:", bool(torch.equal(f_seq, s_seq)), bool(torch.equal(f_val, s_val)),
      "max|dval|", float((f_val - s_val).abs().max()), "val[0:3]", f_val[:3].tolist(), s_val[:3].tolist())
# ---- where do they part?  iteration-0 actions of the fused trace against the staged sampler ----
f_seq2, f_val2, f_key2, tr = opt._plan_raw(x0, keys, seq, sp, trace=True)
print("trace call == plain call:", bool(torch.equal(f_seq2, f_seq)), "key_out equal staged:", bool(torch.equal(f_key.view(torch.int32), s_key.view(torch.int32))))
ks = mbpo_b200.random.split(keys, 2)
carry = ks[:, 0].contiguous()
M = 512 + 15
mean = torch.zeros((B, H, 1), device=dev); std = torch.full((B, H, 1), 0.5, device=dev)
acts = torch.empty((B, M, H, 1), device=dev); nk = torch.empty((B, 2), dtype=torch.uint32, device=dev)
L.check(L.lib.mbpo_icem_sample_actions(L.C.byref(cfg), L.ptr(carry), L.ptr(mean), L.ptr(std), B, L.ptr(acts), L.ptr(nk), None, L.stream_ptr(dev)))
ta = tr["actions"][0].reshape(B, M, H)
eq = (ta == acts.reshape(B, M, H))
print("it0 actions equal frac:", float(eq.float().mean()), "rows fully equal:", int(eq.all(-1).sum()), "of", B * M,
      "first bad rows of problem 0:", (~eq[0].all(-1)).nonzero().flatten()[:10].tolist())
print("fused it0 row0[:5]", ta[0, 0, :5].tolist(), "staged", acts[0, 0, :5, 0].tolist())
vals = torch.empty((B, M), device=dev)
L.check(L.lib.mbpo_rollout_actions(0, L.C.addressof(pp), 0, H, 1, 3, L.ptr(x0), L.ptr(ta.contiguous()), B, M, L.ptr(vals), None, None, None, L.stream_ptr(dev)))
ev = (tr["values"][0] == vals)
print("it0 values equal frac (fused vs staged rollout of the fused actions):", float(ev.float().mean()), "zero-row values", tr["values"][0][0, 512:514].tolist(), vals[0, 512:514].tolist())
