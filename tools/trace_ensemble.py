"""Developer tool: prints the per-step timeline of the ping-pong ensemble rollout kernel (needs a library
built with MBPO_EXTRA_NVCC_FLAGS=-DMBPO_ENS_TRACE)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "model-based-policy-optimizers_b200"))
import runpy  # noqa: E402

runpy.run_path(os.path.join(ROOT, "tools", "profile_ensemble.py"), run_name="__main__")
import mbpo_b200  # noqa: E402

buf = (ctypes.c_longlong * 256)()
mbpo_b200._lib.lib.mbpo_debug_ens_trace(buf)
v = list(buf)
names = {0: "mma: a0(X) ready", 1: "mma: a0(Y) ready"}
T = "XY"
for layer in range(1, 4):
    for tl in range(2):
        names[2 + ((layer - 1) * 2 + tl) * 2] = "mma: L%d(%s) chunks ready" % (layer, T[tl])
        names[3 + ((layer - 1) * 2 + tl) * 2] = "mma: L%d(%s) issued" % (layer, T[tl])
for layer in range(3):
    for tl in range(2):
        st = layer * 2 + tl
        names[20 + st * 6] = "epi w0 : E%d(%s) acc ready" % (layer, T[tl])
        names[60 + st * 6] = "epi w15: E%d(%s) acc ready" % (layer, T[tl])
        for r in range(4):
            names[21 + st * 6 + r] = "epi w0 : E%d(%s) round %d published" % (layer, T[tl], r)
            names[61 + st * 6 + r] = "epi w15: E%d(%s) round %d published" % (layer, T[tl], r)
for tl in range(2):
    names[100 + tl] = "epi w0 : E3(%s) out ready" % T[tl]
    names[102 + tl] = "epi w0 : E3(%s) a0 arrived" % T[tl]
t0 = min(v[i] for i in names if v[i])
ev = sorted((v[i] - t0, names[i]) for i in names if v[i])
prev = 0
for t, n in ev:
    print("%8d  (+%5d)  %s" % (t, t - prev, n))
    prev = t
