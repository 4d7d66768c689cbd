"""Generates tests/golden/replay_golden.npz from the NumPy oracle: jax.random.randint words, a UniformSamplingQueue
insert / sample history (ring wraps), BraxWrapper.reset draws, running_statistics / Normalizer updates, EvalWrapper
metrics and PPO's compute_gae on small seeded inputs (the reference cannot be imported here: no jax, no brax).
Run:  python tests/golden/make_golden_replay.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import brax_replay as br       # noqa: E402
from oracle import jax_prng as jr          # noqa: E402

INSERTS = (5, 20, 7, 32, 3)      # rows per insert into a queue of 32 rows of 10 floats
BATCH = 8


def inputs():
    rng = np.random.default_rng(41)
    rows = [rng.standard_normal((n, 10)).astype(np.float32) for n in INSERTS]
    obs = (rng.standard_normal((6, 50, 3)) * [1.0, 0.3, 8.0] + [0.5, -1.0, 0.0]).astype(np.float32)
    T, B = 12, 9
    gae = dict(reward=rng.standard_normal((T, B)).astype(np.float32),
               values=(3 * rng.standard_normal((T, B))).astype(np.float32),
               bootstrap=rng.standard_normal(B).astype(np.float32),
               truncation=(rng.random((T, B)) < 0.1).astype(np.float32))
    done = np.maximum(gae["truncation"], (rng.random((T, B)) < 0.1).astype(np.float32))
    gae["discount"] = (1 - done).astype(np.float32)
    return rows, obs, gae


def main():
    rows, obs, gae = inputs()
    out = {}
    for legacy in (True, False):
        tag = "legacy" if legacy else "part"
        part = not legacy
        out["randint_%s" % tag] = np.stack([jr.randint(jr.PRNGKey(s), 16, lo, hi, part)
                                            for s, (lo, hi) in enumerate([(0, 10), (-5, 5), (0, 65537), (3, 3)])])
        q = br.UniformSamplingQueue(32, 10, BATCH, part)
        st = q.init(jr.PRNGKey(7))
        for k, r in enumerate(rows):
            st = q.insert(st, r)
            st, batch, idx = q.sample(st)
            out["q_%s_%d_positions" % (tag, k)] = np.array([st.insert_position, st.sample_position])
            out["q_%s_%d_idx" % (tag, k)] = idx
            out["q_%s_%d_batch" % (tag, k)] = batch
            out["q_%s_%d_key" % (tag, k)] = st.key
        out["q_%s_data" % tag] = st.data
        rngs = jr.split(jr.PRNGKey(9), 12, part)
        o, r, k, i = br.brax_wrapper_reset(rngs, br.UniformSamplingQueue(32, 10, 1, part), st, 3, 1)
        out.update({"reset_%s_obs" % tag: o, "reset_%s_reward" % tag: r, "reset_%s_keys" % tag: k,
                    "reset_%s_idx" % tag: i, "reset_%s_rngs" % tag: rngs})
    for k, r in enumerate(rows):
        out["rows_%d" % k] = r
    stats, norm = br.running_statistics_init(3), br.normalizer_init(3)
    for k in range(obs.shape[0]):
        stats = br.running_statistics_update(stats, obs[k], accumulate=np.float64)
        norm = br.normalizer_update(obs[k], norm)
        out["stats_%d" % k] = np.concatenate([[stats["count"]], stats["mean"], stats["summed_variance"], stats["std"]])
        out["norm_%d" % k] = np.concatenate([[norm["size"]], norm["mean"], norm["std"]])
    out["obs"] = obs
    term = ((1 - gae["discount"]) * (1 - gae["truncation"])).astype(np.float32)
    vs, adv = br.compute_gae(gae["truncation"], term, gae["reward"], gae["values"], gae["bootstrap"], 0.95, 0.99)
    out.update({"gae_" + k: v for k, v in gae.items()}, gae_termination=term, gae_vs=vs, gae_advantages=adv)
    er, es, ea = br.eval_metrics(gae["reward"], gae["discount"], np.zeros(9, np.float32), np.zeros(9, np.float32), 1)
    out.update(eval_reward=er, eval_steps=es, eval_active=ea)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "replay_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d arrays, %d bytes)" % (path, len(out), os.path.getsize(path)))


if __name__ == "__main__":
    main()
