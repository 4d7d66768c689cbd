"""CPU tests of the replay-queue oracle (oracle/brax_replay.py, oracle/jax_prng.randint): the restatement of brax's
UniformSamplingQueue and jax.random.randint against independent brute-force statements of the same published
algorithms.  No JAX/brax exists in this image: parity with a real run is unpinned (said so in the oracle headers)."""
import numpy as np
import pytest

from oracle import brax_replay as br
from oracle import jax_prng as jp


def _randint_bigint(key, n, minval, maxval):
    """_randint with python integers: (hi * 2**32 + lo) % span computed the way JAX does, in uint32 steps."""
    k = jp.split(key, 2)
    hi, lo = jp.random_bits(k[0], n), jp.random_bits(k[1], n)
    span = 1 if maxval <= minval else (maxval - minval) % 2 ** 32
    out = []
    for h, l in zip(hi.tolist(), lo.tolist()):
        mult = (2 ** 16) % span
        mult = ((mult * mult) % 2 ** 32) % span
        off = (((h % span) * mult) % 2 ** 32 + l % span) % 2 ** 32
        out.append(minval + off % span)
    return np.array(out, dtype=np.int64)


@pytest.mark.parametrize("minval,maxval", [(0, 10), (3, 3), (7, 2), (-5, 5), (0, 2 ** 31 - 1), (-2 ** 31, 2 ** 31 - 1),
                                            (0, 1), (0, 65536), (0, 65537), (100, 1_000_000)])
def test_randint_matches_bigint_statement(minval, maxval):
    key = jp.PRNGKey(minval * 31 + maxval)
    got = jp.randint(key, 257, minval, maxval)
    want = _randint_bigint(key, 257, minval, maxval)
    assert got.dtype == np.int32 and np.array_equal(got.astype(np.int64), want)
    if maxval > minval:
        assert got.min() >= minval and got.max() < maxval
    else:
        assert np.all(got == minval)


def test_randint_small_spans_are_exactly_the_128_bit_remainder():
    """When the uint32 products cannot wrap (span <= 2**16) the result is (hi * 2**32 + lo) % span."""
    key = jp.PRNGKey(5)
    k = jp.split(key, 2)
    hi, lo = jp.random_bits(k[0], 100).astype(object), jp.random_bits(k[1], 100).astype(object)
    for span in (1, 2, 3, 10, 1000, 65535, 65536):
        got = jp.randint(key, 100, 0, span)
        want = np.array([(int(h) * 2 ** 32 + int(l)) % span for h, l in zip(hi, lo)])
        assert np.array_equal(got, want)


def test_queue_keeps_the_last_rows_in_order():
    """Whatever the insert sizes, the queue holds the most recent min(total, R) rows, oldest first, and
    insert_position / sample_position follow brax's roll arithmetic."""
    rng = np.random.default_rng(0)
    R, D = 37, 9
    q = br.UniformSamplingQueue(R, D, 4)
    st = q.init(jp.PRNGKey(0))
    everything = np.zeros((0, D), np.float32)
    for n in [1, 5, 30, 1, 37, 2, 0, 36, 7, 7, 7]:
        rows = rng.standard_normal((n, D)).astype(np.float32)
        st = q.insert(st, rows)
        everything = np.concatenate([everything, rows])
        live = min(len(everything), R)
        assert st.insert_position == live and st.sample_position == 0
        assert np.array_equal(st.data[:live], everything[len(everything) - live:])
    with pytest.raises(ValueError):
        q.insert(st, np.zeros((R + 1, D), np.float32))


def test_sample_draws_live_rows_and_advances_the_key():
    q = br.UniformSamplingQueue(10, 9, 64)
    st = q.init(jp.PRNGKey(0))
    st0, batch, idx = q.sample(st)                       # empty queue: span 1 -> row 0 (zeros)
    assert np.all(idx == 0) and np.all(batch == 0)
    assert np.array_equal(st0.key, jp.split(jp.PRNGKey(0), 2)[0])
    st = q.insert(st, np.arange(27, dtype=np.float32).reshape(3, 9))
    st1, batch, idx = q.sample(st)
    assert idx.min() >= 0 and idx.max() < 3 and set(idx.tolist()) == {0, 1, 2}
    assert np.array_equal(batch, st.data[idx])
    st2, batch2, idx2 = q.sample(st1)
    assert not np.array_equal(idx, idx2)


def test_brax_wrapper_reset_per_env_keys():
    q = br.UniformSamplingQueue(10, 9, 1)
    st = q.insert(q.init(jp.PRNGKey(0)), np.arange(45, dtype=np.float32).reshape(5, 9))
    rngs = jp.split(jp.PRNGKey(7), 6)
    obs, reward, sys_keys, idx = br.brax_wrapper_reset(rngs, q, st, 3, 1)
    for e in range(6):
        keys = jp.split(rngs[e], 2)
        assert np.array_equal(sys_keys[e], keys[1])
        want = jp.randint(jp.split(keys[0], 2)[1], 1, 0, 5)[0]
        assert idx[e] == want
        assert np.array_equal(obs[e], st.data[want, :3]) and reward[e] == st.data[want, 4]


def test_eval_metrics_fold_matches_a_per_env_statement():
    """EvalWrapper: rewards are summed while the first episode is active; episode_steps freezes at its last step."""
    rng = np.random.default_rng(0)
    T, E, L, rep = 23, 40, 7, 2
    reward = rng.standard_normal((T, E)).astype(np.float32)
    steps0 = (rng.integers(0, 3, E) * rep).astype(np.float32)
    done0 = (rng.random(E) < 0.2).astype(np.float32)
    discount = np.ones((T, E), np.float32)
    want_r, want_s = np.zeros(E, np.float32), np.zeros(E, np.float32)
    for e in range(E):
        s, d, active = steps0[e], done0[e], True
        for t in range(T):
            s = (0 if d else s) + rep
            d = 1.0 if s >= L else 0.0
            discount[t, e] = 1.0 - d
            if active:
                want_r[e] = np.float32(want_r[e] + reward[t, e])
                want_s[e] = s
            active = active and not d
    got_r, got_s, active = br.eval_metrics(reward, discount, steps0, done0, rep)
    assert np.array_equal(got_r, want_r) and np.array_equal(got_s, want_s) and np.all(active == 0)


def test_running_statistics_one_pass_form_equals_the_reference_form():
    """The kernel's single-pass algebra (sum d, sum d*d in float64, one exchange) against the update as brax writes it
    (two passes); also against plain mean / variance of everything seen."""
    rng = np.random.default_rng(0)
    st_a = br.running_statistics_init(3)
    st_b = br.running_statistics_init(3)
    seen = np.zeros((0, 3), np.float32)
    for n in (1, 7, 1000, 50_000):
        batch = (rng.standard_normal((n, 3)) * [1.0, 0.1, 8.0] + [0.5, -2.0, 0.0]).astype(np.float32)
        st_a = br.running_statistics_update(st_a, batch, accumulate=np.float64)
        st_b = br.running_statistics_finalize(st_b, br.running_statistics_sums(batch, st_b["mean"]), n)
        seen = np.concatenate([seen, batch])
        for k in ("count", "mean", "summed_variance", "std"):
            np.testing.assert_allclose(st_b[k], st_a[k], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(st_b["mean"], seen.mean(0, dtype=np.float64), rtol=1e-5, atol=1e-6)
        if len(seen) > 1:
            np.testing.assert_allclose(st_b["std"], seen.std(0, dtype=np.float64), rtol=1e-5, atol=1e-6)
    f32 = br.running_statistics_update(br.running_statistics_init(3), seen)        # float32 sums, as XLA would
    np.testing.assert_allclose(f32["mean"], st_b["mean"], rtol=1e-4, atol=1e-5)
    one = br.running_statistics_update(br.running_statistics_init(2), np.zeros((4, 2), np.float32))
    assert np.all(one["std"] == np.float32(1e-6))                                   # std_min_value clip


def test_bptt_normalizer_matches_plain_mean_and_std():
    """Normalizer.update (bptt_optimizer.py:51-66) is the parallel-variance combine: after any sequence of batches it
    equals mean / std (ddof=0) of everything seen."""
    rng = np.random.default_rng(0)
    st = br.normalizer_init(3)
    seen = np.zeros((0, 3), np.float32)
    for n in (1, 5, 300, 20_000):
        x = (rng.standard_normal((n, 3)) * [1.0, 0.1, 8.0] + [0.5, -2.0, 0.0]).astype(np.float32)
        st = br.normalizer_update(x, st)
        seen = np.concatenate([seen, x])
        assert st["size"] == len(seen)
        np.testing.assert_allclose(st["mean"], seen.mean(0, dtype=np.float64), rtol=2e-5, atol=1e-6)
        if len(seen) > 1:
            np.testing.assert_allclose(st["std"], seen.std(0, dtype=np.float64), rtol=2e-5, atol=1e-6)


def test_compute_gae_reduces_to_discounted_returns():
    """lambda = 1, no truncation / termination: vs_t = sum_k discount^k r_{t+k} + discount^(T-t) bootstrap, and the
    advantage is vs_t - V_t' (one-step form on vs); a termination cuts the bootstrap, a truncation zeroes the delta."""
    rng = np.random.default_rng(0)
    T, B, d = 12, 5, 0.9
    r = rng.standard_normal((T, B)).astype(np.float32)
    v = rng.standard_normal((T, B)).astype(np.float32)
    boot = rng.standard_normal(B).astype(np.float32)
    z = np.zeros((T, B), np.float32)
    vs, adv = br.compute_gae(z, z, r, v, boot, lambda_=1.0, discount=d)
    want = np.zeros((T + 1, B)); want[T] = boot
    for t in range(T - 1, -1, -1):
        want[t] = r[t] + d * want[t + 1]
    np.testing.assert_allclose(vs, want[:T], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(adv, r + d * want[1:] - v, rtol=2e-5, atol=2e-5)
    term = z.copy(); term[7] = 1.0                      # episode ends after step 7: nothing flows back across it
    vs2, _ = br.compute_gae(z, term, r, v, boot, lambda_=1.0, discount=d)
    np.testing.assert_allclose(vs2[7], r[7], rtol=1e-6, atol=1e-6)
    trunc = z.copy(); trunc[3] = 1.0                    # truncated step: delta and advantage masked
    vs3, adv3 = br.compute_gae(trunc, z, r, v, boot, lambda_=0.95, discount=d)
    assert np.all(adv3[3] == 0) and np.array_equal(vs3[3], v[3])


def test_oracle_reproduces_the_replay_golden_vectors():
    """tests/golden/replay_golden.npz (made by make_golden_replay.py) freezes the oracle's answers: integer / byte
    results bit for bit, the float statistics exactly too (same NumPy operations)."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_replay as mg
    g = np.load(os.path.join(here, "golden", "replay_golden.npz"))
    rows, obs, gae = mg.inputs()
    assert np.array_equal(g["obs"], obs) and all(np.array_equal(g["rows_%d" % k], r) for k, r in enumerate(rows))
    for tag, part in (("legacy", False), ("part", True)):
        got = np.stack([jp.randint(jp.PRNGKey(s), 16, lo, hi, part)
                        for s, (lo, hi) in enumerate([(0, 10), (-5, 5), (0, 65537), (3, 3)])])
        assert np.array_equal(got, g["randint_%s" % tag])
        q = br.UniformSamplingQueue(32, 10, mg.BATCH, part)
        st = q.init(jp.PRNGKey(7))
        for k, r in enumerate(rows):
            st = q.insert(st, r)
            st, batch, idx = q.sample(st)
            assert np.array_equal(idx, g["q_%s_%d_idx" % (tag, k)]) and np.array_equal(batch, g["q_%s_%d_batch" % (tag, k)])
            assert np.array_equal(st.key, g["q_%s_%d_key" % (tag, k)])
            assert [st.insert_position, st.sample_position] == g["q_%s_%d_positions" % (tag, k)].tolist()
        assert np.array_equal(st.data, g["q_%s_data" % tag])
    vs, adv = br.compute_gae(gae["truncation"], g["gae_termination"], gae["reward"], gae["values"], gae["bootstrap"],
                             0.95, 0.99)
    assert np.array_equal(vs, g["gae_vs"]) and np.array_equal(adv, g["gae_advantages"])
