"""CPU tests of the oracle itself: the NumPy restatement against its independent plain-C twin,
against the committed golden vectors, and against the reference's own behavioural tests
(tests/test_sys_pendulum.py shapes; tests/test_icemopt.py threshold sum(rewards) >= -400)."""
import os

import numpy as np
import pytest

from oracle import c_twin
from oracle import jax_prng as jr
from oracle import mbpo_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "icem_golden.npz"))
P9 = orc.PendulumParams().packed()


def _states(n, seed):
    rng = np.random.default_rng(seed)
    th, w = rng.uniform(-np.pi, np.pi, n), rng.uniform(-8, 8, n)
    return np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)


@pytest.fixture(scope="module")
def twin():
    return c_twin.load()


# ---- golden vectors (frozen oracle outputs; tests/golden/make_golden.py) -------------------------
def test_golden_prng_words_bit_exact():
    key = jr.PRNGKey(1234)
    for part in (0, 1):
        assert np.array_equal(jr.split(key, 5, bool(part)), GOLD["split5_p%d" % part])
        assert np.array_equal(jr.random_bits(key, 11, bool(part)), GOLD["bits11_p%d" % part])
        np.testing.assert_allclose(jr.normal(key, 16, bool(part)), GOLD["normal16_p%d" % part], rtol=1e-6)


def test_golden_noise_step_rollout_plan():
    keys = GOLD["noise_keys"]
    for h, ex in ((20, 0.0), (30, 2.0), (15, 1.0)):
        np.testing.assert_allclose(orc.powerlaw_psd_gaussian_keys(ex, h, keys), GOLD["noise_h%d_e%g" % (h, ex)],
                                   rtol=1e-6, atol=1e-6)
    xn, r = orc.pendulum_step(GOLD["step_x"], GOLD["step_u"])
    np.testing.assert_allclose(xn, GOLD["step_xn"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(r, GOLD["step_r"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.rollout_actions(GOLD["step_x"], GOLD["roll_actions"]), GOLD["roll_returns"],
                               rtol=1e-5)
    p = orc.ICemParams(num_samples=64, num_elites=8, num_particles=1, num_steps=3)
    st = orc.ICemState(key=GOLD["plan_key_in"], best_sequence=np.zeros((20, 1), np.float32), best_reward=np.float32(0))
    trace = []
    new = orc.icem_optimize(GOLD["plan_x0"], st, p, 20, trace=trace)
    assert np.array_equal(new.key, GOLD["plan_key_out"])
    for i, t in enumerate(trace):
        assert np.array_equal(t["elite_idx"], GOLD["plan_it%d_elite_idx" % i])
        np.testing.assert_allclose(t["actions"], GOLD["plan_it%d_actions" % i], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(t["mean"], GOLD["plan_it%d_mean" % i], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(new.best_sequence, GOLD["plan_best_seq"], rtol=1e-6, atol=1e-6)


# ---- the reference's own tests, restated ------------------------------------------------------------
def test_sys_pendulum_shapes():
    """tests/test_sys_pendulum.py: vmap(reset) over 20 keys, vmap(step) with uniform actions."""
    keys = jr.split(jr.PRNGKey(0), 21)
    x = np.stack([orc.pendulum_reset()[0] for _ in keys[1:]])
    action_key = jr.split(keys[0], 2)[0]
    actions = jr.uniform(action_key, 20)
    xn, r = orc.pendulum_step(x, actions)
    assert xn.shape == (20, 3) and r.shape == (20,)
    assert np.allclose(x, [-1, 0, 0])


@pytest.mark.parametrize("partitionable", [False, True])
def test_icemopt_threshold_c_twin(twin, partitionable):
    """tests/test_icemopt.py:37-38 with the C twin (fast): sum of 200 closed-loop rewards >= -400."""
    key = jr.split(jr.PRNGKey(0), 3, partitionable)[1]
    st = orc.icem_init(key, 20, 1, partitionable)
    cfg = c_twin.make_cfg(orc.ICemParams(num_particles=1), 20, partitionable)
    states, rewards, actions, seq, k = c_twin.closed_loop(twin, cfg, P9, np.array([-1, 0, 0], np.float32), st.key,
                                                          st.best_sequence, 200)
    assert rewards.sum() >= -400, rewards.sum()
    assert abs(rewards[-1]) <= 0.5 and states[-1, 0] > 0.95          # upright at the end


def test_icemopt_threshold_numpy_oracle():
    xs, rs, us = orc.closed_loop_mpc(200, 20)
    assert rs.sum() >= -400, rs.sum()
    assert rs.sum() == pytest.approx(float(GOLD["mpc_return200"]), abs=1e-3)
    np.testing.assert_allclose(us[:10], GOLD["mpc_actions10"], rtol=1e-5, atol=1e-6)


def test_particles_are_identical_for_the_deterministic_pendulum():
    """the pendulum ignores the particle key, so P=10 only averages P equal values."""
    p10 = orc.ICemParams(num_samples=40, num_elites=6, num_steps=2, num_particles=10)
    p1 = orc.ICemParams(num_samples=40, num_elites=6, num_steps=2, num_particles=1)
    st = orc.icem_init(jr.PRNGKey(3), 8)
    x0 = _states(1, 0)[0]
    a = orc.icem_optimize(x0, st, p10, 8)
    b = orc.icem_optimize(x0, st, p1, 8)
    np.testing.assert_allclose(a.best_reward, b.best_reward, rtol=1e-6)
    assert np.array_equal(a.key, b.key)


def test_prev_elites_closure_quirk_and_mean_objective():
    """kept-elite rows are all zero in every iteration; the
    objective is the horizon-mean reward."""
    p = orc.ICemParams(num_samples=30, num_elites=5, num_steps=3, num_particles=1)
    st = orc.icem_init(jr.PRNGKey(4), 8)
    trace = []
    x0 = _states(1, 1)[0]
    orc.icem_optimize(x0, st, p, 8, trace=trace)
    for t in trace:
        assert t["actions"].shape == (30 + p.num_prev_elites, 8, 1)
        assert np.all(t["actions"][30:] == 0)
        _, _, rew, _ = orc.rollout_actions(np.repeat(x0[None], 31, 0), t["actions"][:, :, 0], full=True)
        np.testing.assert_allclose(t["values"], rew.mean(axis=1), rtol=2e-6)
    assert p.num_prev_elites == 1 and orc.ICemParams().num_prev_elites == 15


def test_stable_argsort_total_order():
    v = np.array([1.0, -0.0, 0.0, np.nan, -np.inf, 1.0, np.inf, -3.0], np.float32)
    assert orc.stable_argsort(v).tolist() == [4, 7, 1, 2, 0, 5, 6, 3]          # -0 == +0 tie by index; NaN last


# ---- NumPy oracle vs the plain-C twin ---------------------------------------------------------------
@pytest.mark.parametrize("horizon,exponent", [(5, 0.0), (8, 1.0), (15, 2.0), (20, 0.0), (30, 2.0), (50, 0.5)])
@pytest.mark.parametrize("partitionable", [False, True])
def test_twin_powerlaw_noise(twin, horizon, exponent, partitionable):
    import ctypes as C
    keys = jr.split(jr.PRNGKey(horizon), 50, partitionable)
    out = np.empty((50, horizon), np.float32)
    twin.orc_powerlaw_noise(keys.ctypes.data_as(C.c_void_p), 50, horizon, C.c_float(exponent), int(partitionable),
                            out.ctypes.data_as(C.c_void_p))
    want = orc.powerlaw_psd_gaussian_keys(exponent, horizon, keys, partitionable)
    np.testing.assert_allclose(out, want, rtol=1e-5, atol=5e-6)
    assert 0.7 < want.std() < 1.4                                               # ~unit-variance normalisation


def test_twin_pendulum_and_rollouts(twin):
    import ctypes as C
    x = _states(500, 2)
    u = np.random.default_rng(3).uniform(-1.5, 1.5, 500).astype(np.float32)
    xn, r = np.empty((500, 3), np.float32), np.empty(500, np.float32)
    f = lambda a: a.ctypes.data_as(C.c_void_p)
    twin.orc_pendulum_step(f(P9), f(x), f(u), 500, f(xn), f(r))
    wxn, wr = orc.pendulum_step(x, u)
    np.testing.assert_allclose(xn, wxn, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(r, wr, rtol=1e-5, atol=2e-6)
    acts = np.clip(np.random.default_rng(4).normal(0, .5, (6, 40, 20)), -1, 1).astype(np.float32)
    ret = np.empty((6, 40), np.float32)
    twin.orc_rollout_returns(f(P9), f(x[:6].copy()), f(acts), 6, 40, 20, f(ret))
    want = orc.rollout_actions(np.repeat(x[:6], 40, 0), acts.reshape(240, 20)).reshape(6, 40)
    bad = np.abs(ret - want) > 1e-5 * np.abs(want) + 1e-6
    assert bad.mean() <= 0.01


@pytest.mark.parametrize("params", [dict(num_samples=64, num_elites=8, num_particles=1, num_steps=3),
                                    dict(num_samples=100, num_elites=10, num_particles=10, num_steps=2, alpha=0.1,
                                         exponent=2.0)])
def test_twin_icem_optimize(twin, params):
    p = orc.ICemParams(**params)
    B, H = 6, 20
    x0 = _states(B, 5)
    keys = jr.split(jr.PRNGKey(6), B)
    seq = np.random.default_rng(7).uniform(-1, 1, (B, H)).astype(np.float32)
    o_seq, o_val, o_key, used = c_twin.optimize_batch(twin, c_twin.make_cfg(p, H), P9, x0, keys, seq, num_threads=2)
    agree = 0
    for b in range(B):
        st = orc.ICemState(key=keys[b], best_sequence=seq[b][:, None], best_reward=np.float32(0))
        new = orc.icem_optimize(x0[b], st, p, H)
        assert np.array_equal(o_key[b], new.key)                                 # integer path: bit-exact
        if np.allclose(o_seq[b], new.best_sequence[:, 0], rtol=1e-5, atol=5e-6):
            agree += 1
            assert o_val[b] == pytest.approx(float(new.best_reward), rel=1e-5, abs=1e-6)
    assert agree >= B - 1          # an elite flip at a sub-tolerance return gap may change one problem


def test_twin_env_rollout(twin):
    x0 = GOLD["env_x0"]
    acts = GOLD["env_actions"]
    out = c_twin.env_rollout(twin, P9, x0, acts, episode_length=10)
    assert np.array_equal(out["discount"], GOLD["env_discount"])
    assert np.array_equal(out["truncation"], GOLD["env_truncation"])
    # teacher-forced float parity (open-loop trajectories diverge chaotically)
    xn, r = orc.pendulum_step(out["observation"].reshape(-1, 3), acts.reshape(-1))
    done = (1 - out["discount"]).reshape(-1, 1)
    nxt = np.where(done != 0, np.broadcast_to(x0[None], out["observation"].shape).reshape(-1, 3), xn)
    np.testing.assert_allclose(out["next_observation"].reshape(-1, 3), nxt, rtol=1e-5, atol=3e-6)
    np.testing.assert_allclose(out["reward"].reshape(-1), r, rtol=1e-5, atol=3e-6)
    want = orc.env_rollout(x0, acts, 10)
    assert np.array_equal(want["discount"], GOLD["env_discount"]) and np.array_equal(want["final_steps"],
                                                                                      out["final_steps"])
    # episode bookkeeping: done every 10 steps, truncation == done (system done is always 0)
    assert np.array_equal(out["truncation"], 1 - out["discount"])
    assert np.all(out["discount"][9::10] == 0) and out["discount"].sum() == acts.size - (25 // 10) * 40


def test_mlp_oracle_bf16_rounding():
    ens = orc.make_mlp_ensemble(seed=3, members=5)
    inp = np.random.default_rng(0).uniform(-1, 1, (64, 4)).astype(np.float32)
    member = np.arange(64) % 5
    a = orc.mlp_member_forward(ens, member, inp, bf16=False)
    b = orc.mlp_member_forward(ens, member, inp, bf16=True)
    assert a.shape == (64, 3) and np.abs(a - b).max() < 5e-2 and np.abs(a - b).max() > 0
    r = orc._bf16_round(np.array([1.0, 1.00390625, 1.0058594, -2.5], np.float32))
    assert r.tolist() == [1.0, 1.0, 1.0078125, -2.5]                              # round to nearest even


def test_oracle_cost_fn_penalty_and_array_bounds():
    """icem_optimizer.py:161-166 (reward - lambda * relu(cost), mean / max over identical particles) and
    :47-48,191 (bounds broadcastable to (H, A)) in the oracle."""
    horizon = 12
    p = orc.ICemParams(num_samples=64, num_elites=8, num_particles=3, num_steps=2, lambda_constraint=10.0)
    x0 = np.array([np.cos(2.0), np.sin(2.0), 6.5], np.float32)
    key = jr.PRNGKey(5)
    mean = np.zeros((horizon, 1), np.float32)
    std = np.full((horizon, 1), 0.5, np.float32)
    _, acts, _ = orc.icem_sample_actions(key, mean, std, p, horizon)
    plain = orc.icem_objective(x0, acts, p, orc.PendulumParams())
    never = lambda obs, a: (np.abs(obs[:, :, 2]).max(axis=1) - np.float32(1e6)).astype(np.float32)
    binds = lambda obs, a: (np.abs(obs[:, :, 2]).max(axis=1) - np.float32(5.0)).astype(np.float32)
    assert np.array_equal(orc.icem_objective(x0, acts, p, orc.PendulumParams(), cost_fn=never), plain)
    pen = orc.icem_objective(x0, acts, p, orc.PendulumParams(), cost_fn=binds)
    pes = orc.icem_objective(x0, acts, p, orc.PendulumParams(), cost_fn=binds, use_pessimism=True)
    assert np.all(pen < plain) and np.all(pes < plain)                   # thdot0 = 6.5 > 5: the constraint binds
    c = binds(orc.rollout_actions(x0, acts[:, :, 0], full=True)[1], acts)
    np.testing.assert_allclose(plain - pes, 10.0 * c, rtol=1e-6)
    np.testing.assert_allclose(pen, pes, rtol=1e-5, atol=1e-5)            # mean of 3 identical vs max: rounding only
    # array-valued bounds
    u_min = -np.linspace(0.1, 1.0, horizon, dtype=np.float32).reshape(horizon, 1)
    u_max = np.linspace(1.0, 0.2, horizon, dtype=np.float32).reshape(horizon, 1)
    pb = orc.ICemParams(num_samples=64, num_elites=8, num_particles=1, num_steps=2, u_min=u_min, u_max=u_max)
    _, ab, _ = orc.icem_sample_actions(key, mean, std, pb, horizon)
    assert np.all(ab >= u_min[None]) and np.all(ab <= u_max[None])
    assert np.array_equal(ab[:64], np.clip(acts[:64] * 0 + (mean[None] + (acts[:64] - acts[:64])) + ab[:64], u_min, u_max))
    st = orc.icem_optimize(x0, orc.icem_init(jr.PRNGKey(1), horizon), pb, horizon)
    assert np.all(st.best_sequence >= u_min) and np.all(st.best_sequence <= u_max)


def test_oracle_actor_rollout_conventions():
    """Policy-in-the-loop oracle: key conventions of sac.py:288-292 vs acting.py:68-73, NormalTanh bounds,
    deterministic mode, and agreement with env_rollout on the sampled actions."""
    pol = orc.make_policy_params(seed=7)
    E, T, L = 40, 9, 4
    rng = np.random.default_rng(3)
    th, w = rng.uniform(-np.pi, np.pi, E), rng.uniform(-8, 8, E)
    x0 = np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)
    key = jr.PRNGKey(2)
    sac, k_sac = orc.actor_rollout(pol, x0, key, T, L, key_convention="sac")
    unr, k_unr = orc.actor_rollout(pol, x0, key, T, L, key_convention="unroll")
    ks = jr.split(key, 2)
    # first step: sac samples with the SECOND half of split(key), generate_unroll with the FIRST
    assert np.array_equal(sac["action"][0], orc.policy_sample(pol, x0, ks[1]))
    assert np.array_equal(unr["action"][0], orc.policy_sample(pol, x0, ks[0]))
    assert not np.array_equal(k_sac, k_unr)
    k = key
    for _ in range(T):
        k = jr.split(k, 2)[0]
    assert np.array_equal(k, k_sac)
    assert np.all(np.abs(sac["action"]) <= 1.0)
    env = orc.env_rollout(x0, sac["action"][..., 0], L)
    for f in ("reward", "discount", "next_observation", "truncation", "observation"):
        assert np.array_equal(env[f], sac[f]), f
    assert sac["truncation"][L - 1].all() and not sac["truncation"][L - 2].any()
    det, _ = orc.actor_rollout(pol, x0, key, 2, L, deterministic=True)
    assert np.array_equal(det["action"][0], np.tanh(orc.policy_logits(pol, x0)[:, :1]).astype(np.float32))
    one, k1 = orc.actor_rollout(pol, x0, key, 1, L, key_convention="as_is")
    assert np.array_equal(k1, key) and np.array_equal(one["action"][0], orc.policy_sample(pol, x0, key))
