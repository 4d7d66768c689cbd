"""Generates tests/golden/bptt_golden.npz from the NumPy oracle: rollout_policy with BPTT's actor, its
cotangent pass and lambda_return on small seeded inputs (the reference cannot be imported here: no jax).
Run:  python tests/golden/make_golden_bptt.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import jax_prng as jr          # noqa: E402
from oracle import mbpo_oracle as orc      # noqa: E402


def inputs():
    rng = np.random.default_rng(31)
    th, w = rng.uniform(-np.pi, np.pi, 16), rng.uniform(-8, 8, 16)
    x0 = np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)
    actor = orc.BpttActorParams(mlp=orc.make_policy_params(seed=33, hidden=(64, 64)), init_stddev=0.5,
                                obs_mean=np.array([0.1, -0.2, 0.5], np.float32),
                                obs_std=np.array([0.7, 0.8, 3.0], np.float32))
    g = dict(g_reward=rng.standard_normal((16, 10)).astype(np.float32),
             g_next_obs=rng.standard_normal((16, 10, 3)).astype(np.float32),
             g_obs=rng.standard_normal((16, 10, 3)).astype(np.float32),
             g_action=rng.standard_normal((16, 10, 1)).astype(np.float32),
             next_values=rng.standard_normal((16, 10)).astype(np.float32))
    return x0, actor, jr.PRNGKey(35), g


def main():
    x0, actor, key, g = inputs()
    out = dict(x0=x0, key=key, **g)
    tr, key_out = orc.rollout_policy(actor, x0, key, 10)
    out.update({"tr_" + k: v for k, v in tr.items()}, key_out=key_out)
    ga, gx0 = orc.rollout_policy_vjp(tr["observation"], tr["action"], g["g_reward"], g["g_next_obs"], g["g_obs"],
                                     g["g_action"])
    out.update(vjp_g_action=ga, vjp_g_x0=gx0)
    out["lambda_returns"] = orc.lambda_return(tr["reward"], g["next_values"], 0.99, 0.95)
    gr, gnv = orc.lambda_return_vjp(g["g_reward"], 0.99, 0.95)
    out.update(lambda_g_reward=gr, lambda_g_next_values=gnv)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bptt_golden.npz"), **out)
    print("wrote bptt_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
