"""Generates tests/golden/icem_golden.npz from the NumPy oracle (the reference cannot be imported
here: no jax).  The fixture freezes the oracle's outputs on small seeded inputs so that (a) a
later edit of the oracle cannot silently move the target and (b) the GPU tests have committed
vectors to compare against.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import jax_prng as jr          # noqa: E402
from oracle import mbpo_oracle as orc      # noqa: E402


def states(n, seed):
    rng = np.random.default_rng(seed)
    th, w = rng.uniform(-np.pi, np.pi, n), rng.uniform(-8, 8, n)
    return np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)


def main():
    out = {}
    # PRNG words (bit-exact targets)
    key = jr.PRNGKey(1234)
    for part in (0, 1):
        out["split5_p%d" % part] = jr.split(key, 5, bool(part))
        out["bits11_p%d" % part] = jr.random_bits(key, 11, bool(part))
        out["normal16_p%d" % part] = jr.normal(key, 16, bool(part))
    # colored noise rows
    keys = jr.split(jr.PRNGKey(7), 6)
    out["noise_keys"] = keys
    for h, ex in ((20, 0.0), (30, 2.0), (15, 1.0)):
        out["noise_h%d_e%g" % (h, ex)] = orc.powerlaw_psd_gaussian_keys(ex, h, keys)
    # pendulum steps and rollouts
    x = states(64, 1)
    u = np.random.default_rng(2).uniform(-1.2, 1.2, 64).astype(np.float32)
    xn, r = orc.pendulum_step(x, u)
    out.update(step_x=x, step_u=u, step_xn=xn, step_r=r)
    acts = np.clip(np.random.default_rng(3).normal(0, 0.5, (64, 20)), -1, 1).astype(np.float32)
    out.update(roll_actions=acts, roll_returns=orc.rollout_actions(x, acts))
    # one full iCEM plan with per-iteration trace, small population
    p = orc.ICemParams(num_samples=64, num_elites=8, num_particles=1, num_steps=3)
    st = orc.icem_init(jr.PRNGKey(5), 20)
    trace = []
    new = orc.icem_optimize(x[0], st, p, 20, trace=trace)
    out.update(plan_x0=x[0], plan_key_in=st.key, plan_key_out=new.key, plan_best_seq=new.best_sequence,
               plan_best_reward=new.best_reward)
    for i, t in enumerate(trace):
        out["plan_it%d_actions" % i] = t["actions"]
        out["plan_it%d_values" % i] = t["values"]
        out["plan_it%d_elite_idx" % i] = t["elite_idx"].astype(np.int32)
        out["plan_it%d_mean" % i] = t["mean"]
        out["plan_it%d_std" % i] = t["std"]
    # config 1 closed loop (tests/test_icemopt.py), first 10 steps + the 200-step return
    xs, rs, us = orc.closed_loop_mpc(200, 20)
    out.update(mpc_states10=xs[:10], mpc_rewards10=rs[:10], mpc_actions10=us[:10], mpc_return200=np.float32(rs.sum()))
    # wrapped env rollout
    e_x0 = states(40, 4)
    e_act = np.random.default_rng(5).uniform(-1, 1, (25, 40)).astype(np.float32)
    env = orc.env_rollout(e_x0, e_act, episode_length=10)
    out.update(env_x0=e_x0, env_actions=e_act, env_next_obs=env["next_observation"], env_reward=env["reward"],
               env_discount=env["discount"], env_truncation=env["truncation"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "icem_golden.npz"), **out)
    print("wrote icem_golden.npz with %d arrays" % len(out))


if __name__ == "__main__":
    main()
