"""Regenerates tests/golden/jax_golden.npz from the REAL reference (mbpo + JAX), the day a JAX install is at hand.

This container and the GPU boxes have no jax / brax / flax / distrax wheels, so every float of the
path is currently checked against the NumPy oracle and a float64 error budget, and only the integer PRNG path is
pinned to external known answers.  With JAX available,

    pip install "jax[cpu]==0.4.30" brax flax optax distrax chex jaxtyping
    PYTHONPATH=/path/to/Model-based-policy-optimizers python tests/golden/make_golden_jax.py

writes outputs of the reference's own functions on the seeded inputs below; tests/test_jax_golden.py then checks the
oracle (CPU) and the CUDA path (GPU) against them with the north star's tolerances (uint32 words and keys bit-exact,
floats rel 1e-5) and is skipped while the file is absent.  Nothing in the product imports this script.

Each array names the reference call that produced it.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def states(n, seed):
    rng = np.random.default_rng(seed)
    th, w = rng.uniform(-np.pi, np.pi, n), rng.uniform(-8, 8, n)
    return np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)


def main():
    import jax
    import jax.numpy as jnp
    import jax.random as jr
    from mbpo.optimizers import iCemParams, iCemTO                     # mbpo/optimizers/__init__.py:1-6
    from mbpo.systems import PendulumSystem                            # mbpo/systems/__init__.py:1-4
    from mbpo.utils.general_utils import powerlaw_psd_gaussian         # general_utils.py:81-208
    from mbpo.utils.optimizer_utils import rollout_actions             # optimizer_utils.py:11-59

    out = {"jax_version": np.array(jax.__version__), "threefry_partitionable":
           np.array(bool(jax.config.jax_threefry_partitionable))}
    u32 = lambda k: np.asarray(jr.key_data(k) if hasattr(jr, "key_data") else k, dtype=np.uint32)

    # ---- PRNG (jax.random; call sites icem_optimizer.py:123,155,174-180,246, general_utils.py:189-191) ----------
    key = jr.PRNGKey(1234)
    out["prngkey_1234"] = u32(key)
    out["prngkey_minus1"] = u32(jr.PRNGKey(-1))
    out["split5"] = u32(jr.split(key, 5))
    out["bits11"] = np.asarray(jr.bits(key, (11,), dtype=jnp.uint32))
    out["normal16"] = np.asarray(jr.normal(key, (16,)))
    out["uniform9"] = np.asarray(jr.uniform(key, (9,), minval=-2.0, maxval=3.0))
    out["randint33_0_10"] = np.asarray(jr.randint(key, (33,), 0, 10))
    out["randint7_5_1000000"] = np.asarray(jr.randint(key, (7,), 5, 1000000))

    # ---- powerlaw_psd_gaussian (general_utils.py:81-208) ------------------------------------------------------
    nkeys = jr.split(jr.PRNGKey(7), 6)
    out["noise_keys"] = u32(nkeys)
    for h, ex in ((20, 0.0), (30, 2.0), (15, 1.0), (25, 0.5), (50, 1.0)):
        out["noise_h%d_e%g" % (h, ex)] = np.stack([np.asarray(powerlaw_psd_gaussian(ex, h, k)) for k in nkeys])

    # ---- PendulumSystem.step (pendulum_system.py:18-39) and rollout_actions ------------------------------------
    system = PendulumSystem()
    sp = system.reset(jr.PRNGKey(0)).system_params
    x = states(64, 1)
    u = np.random.default_rng(2).uniform(-1.2, 1.2, (64, 1)).astype(np.float32)
    st = jax.vmap(lambda xx, uu: system.step(xx, uu, sp))(jnp.asarray(x), jnp.asarray(u))
    out.update(step_x=x, step_u=u, step_xn=np.asarray(st.x_next), step_r=np.asarray(st.reward))
    for h in (20, 30, 50):
        x0 = states(16, 10 + h)
        acts = np.clip(np.random.default_rng(h).normal(0, 0.5, (16, h, 1)), -1, 1).astype(np.float32)
        tr = jax.vmap(lambda xx, aa: rollout_actions(system, sp, xx, aa, h))(jnp.asarray(x0), jnp.asarray(acts))
        out.update({"roll_h%d_x0" % h: x0, "roll_h%d_actions" % h: acts, "roll_h%d_reward" % h: np.asarray(tr.reward),
                    "roll_h%d_observation" % h: np.asarray(tr.observation),
                    "roll_h%d_next_observation" % h: np.asarray(tr.next_observation)})

    # ---- iCemTO.init / optimize / act (icem_optimizer.py:121-257), reference defaults and config 2's shape ------
    for tag, h, params in (("defaults_h20", 20, iCemParams()),
                           ("config2_h30", 30, iCemParams(num_samples=512, num_particles=1)),
                           ("colored_h30", 30, iCemParams(num_samples=512, num_particles=1, exponent=2.0, alpha=0.1))):
        opt = iCemTO(horizon=h, action_dim=1, system=None, opt_params=params, key=jr.PRNGKey(5))
        opt.set_system(system)
        state = opt.init(jr.PRNGKey(11))
        x0 = states(4, 20 + h)
        seqs, vals, keys_out, first = [], [], [], []
        for i in range(4):
            new = opt.optimize(jnp.asarray(x0[i]), state)
            seqs.append(np.asarray(new.best_sequence)); vals.append(np.asarray(new.best_reward))
            keys_out.append(u32(new.key)); first.append(np.asarray(new.action))
        out.update({"icem_%s_x0" % tag: x0, "icem_%s_state_key" % tag: u32(state.key),
                    "icem_%s_best_sequence" % tag: np.stack(seqs), "icem_%s_best_reward" % tag: np.stack(vals),
                    "icem_%s_key" % tag: np.stack(keys_out), "icem_%s_action" % tag: np.stack(first)})

    # ---- the reference's own test (tests/test_icemopt.py): 200-step closed loop -------------------------------
    key = jr.PRNGKey(0)
    optimizer_key, init_key, key = jr.split(key, 3)
    sys_state = system.reset(key)
    cem = iCemTO(horizon=20, action_dim=1, system=None, opt_params=iCemParams(), key=optimizer_key)
    cem.set_system(system)
    cst = cem.init(init_key)

    def body(carry, _):
        s, c = carry
        a, c2 = cem.act(obs=s.x_next, opt_state=c)
        s2 = system.step(x=s.x_next, u=a, system_params=s.system_params)
        c2 = c2.replace(system_params=s2.system_params)
        return [s2, c2], [s2.x_next, s2.reward, a]

    _, (xs, rs, us) = jax.lax.scan(body, [sys_state, cst], xs=None, length=200)
    out.update(mpc_states=np.asarray(xs), mpc_rewards=np.asarray(rs), mpc_actions=np.asarray(us))

    path = os.path.join(HERE, "jax_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, jax %s, sum(rewards) of the closed loop = %.3f)" % (
        path, len(out), jax.__version__, float(np.asarray(rs).sum())))


if __name__ == "__main__":
    try:
        import jax  # noqa: F401
    except ImportError:
        sys.exit("make_golden_jax.py needs jax (and the reference's other dependencies): none is installed here -- "
                 "see the docstring")
    main()
