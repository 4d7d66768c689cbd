"""Developer tool: prints the per-step timeline of the ensemble rollout kernel (needs a library
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
rc = mbpo_b200._lib.lib.mbpo_debug_ens_trace(buf)
v = list(buf)
t0 = v[0]
names = {0: "mma: a0 ready", 1: "mma: mma0 committed", 56: "epi w0: bar_out passed", 57: "epi w0: a0 arrived"}
for layer in range(3):
    for r in range(4):
        names[2 + layer * 8 + r * 2] = "mma: L%d chunk %d ready" % (layer + 1, r)
        names[3 + layer * 8 + r * 2] = "mma: L%d chunk %d issued" % (layer + 1, r)
        names[33 + layer * 8 + r] = "epi w0: L%d round %d arrived" % (layer, r)
        names[65 + layer * 8 + r] = "epi w15: L%d round %d arrived" % (layer, r)
    names[32 + layer * 8] = "epi w0: L%d bar_mma passed" % layer
    names[64 + layer * 8] = "epi w15: L%d bar_mma passed" % layer
ev = sorted((v[i] - t0, names[i]) for i in names if v[i])
prev = 0
for t, n in ev:
    print("%8d  (+%5d)  %s" % (t, t - prev, n))
    prev = t
