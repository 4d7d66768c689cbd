"""CPU tests of the BPTT part of the oracle (rollout_policy, its cotangent pass, lambda_return): against the
committed golden vectors, against central differences of the oracle's own forward functions in float64, and
against torch autograd through a torch restatement of the same computation (an independent derivation)."""
import os

import numpy as np
import torch

from oracle import jax_prng as jr
from oracle import mbpo_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "bptt_golden.npz"))


def _golden_actor():
    return orc.BpttActorParams(mlp=orc.make_policy_params(seed=33, hidden=(64, 64)), init_stddev=0.5,
                               obs_mean=np.array([0.1, -0.2, 0.5], np.float32),
                               obs_std=np.array([0.7, 0.8, 3.0], np.float32))


def test_golden_rollout_policy_vjp_lambda():
    tr, key_out = orc.rollout_policy(_golden_actor(), GOLD["x0"], GOLD["key"], 10)
    assert np.array_equal(key_out, GOLD["key_out"])
    for k, v in tr.items():
        np.testing.assert_allclose(v, GOLD["tr_" + k], rtol=1e-6, atol=1e-7, err_msg=k)
    ga, gx0 = orc.rollout_policy_vjp(GOLD["tr_observation"], GOLD["tr_action"], GOLD["g_reward"], GOLD["g_next_obs"],
                                     GOLD["g_obs"], GOLD["g_action"])
    np.testing.assert_allclose(ga, GOLD["vjp_g_action"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gx0, GOLD["vjp_g_x0"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(orc.lambda_return(GOLD["tr_reward"], GOLD["next_values"], 0.99, 0.95), GOLD["lambda_returns"])
    gr, gnv = orc.lambda_return_vjp(GOLD["g_reward"], 0.99, 0.95)
    np.testing.assert_allclose(gr, GOLD["lambda_g_reward"], rtol=1e-6)
    np.testing.assert_allclose(gnv, GOLD["lambda_g_next_values"], rtol=1e-6)


def test_rollout_policy_semantics():
    """Shared draw per step (the key is not vmapped, bptt_optimizer.py:366-368), squashing into +-0.999, evaluate
    leaves the key alone, observation[t+1] = next_observation[t]."""
    actor = _golden_actor()
    x0 = GOLD["x0"]
    tr, key_out = orc.rollout_policy(actor, x0, GOLD["key"], 6)
    one, _ = orc.rollout_policy(actor, x0[3:4], GOLD["key"], 6)
    assert np.array_equal(one["action"][0], tr["action"][3])
    assert np.abs(tr["action"]).max() <= np.float32(0.999)
    assert np.array_equal(tr["observation"][:, 1:], tr["next_observation"][:, :-1]) and np.array_equal(tr["observation"][:, 0], x0)
    k = GOLD["key"]
    for _ in range(6):
        k = jr.split(k, 2)[1]
    assert np.array_equal(key_out, k)
    ev, key_ev = orc.rollout_policy(actor, x0, GOLD["key"], 6, evaluate=True)
    assert np.array_equal(key_ev, GOLD["key"])
    mu, _ = orc.bptt_actor(actor, x0)
    np.testing.assert_allclose(ev["action"][:, 0], np.clip(np.tanh(mu), -0.999, 0.999), rtol=1e-6)


def test_pendulum_step_vjp_central_differences():
    rng = np.random.default_rng(0)
    n = 256
    th = rng.uniform(-np.pi, np.pi, n)
    x = np.stack([np.cos(th), np.sin(th), rng.uniform(-7, 7, n)], -1)
    u = rng.uniform(-0.99, 0.99, n)
    gn, gr = rng.standard_normal((n, 3)), rng.standard_normal(n)
    gx, gu = orc.pendulum_step_vjp(x, u, gn, gr, dtype=np.float64)

    def loss(x_, u_):
        nx, r = orc.pendulum_step(x_, u_, dtype=np.float64)
        return (nx * gn).sum(-1) + r * gr
    eps = 1e-6
    for i in range(3):
        d = np.zeros_like(x)
        d[:, i] = eps
        np.testing.assert_allclose((loss(x + d, u) - loss(x - d, u)) / (2 * eps), gx[:, i], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose((loss(x, u + eps) - loss(x, u - eps)) / (2 * eps), gu, rtol=1e-5, atol=1e-6)
    # saturated torque / speed: jnp.clip passes no cotangent outside its interval, half of it at a tie
    assert orc._clip_grad(np.array([-2.0, -1.0, 0.0, 1.0, 2.0]), -1.0, 1.0).tolist() == [0.0, 0.5, 1.0, 0.5, 0.0]
    _, gu_sat = orc.pendulum_step_vjp(x, np.full(n, 1.5), gn, np.zeros(n), dtype=np.float64)
    assert np.all(gu_sat == 0)


def test_rollout_vjp_and_lambda_return_match_torch_autograd():
    """An independent derivation: torch autograd through a float64 torch restatement of the open-loop rollout
    (actions as leaves = the stop_gradient structure) and of lambda_return."""
    rng = np.random.default_rng(5)
    B, H = 8, 9
    th = rng.uniform(-np.pi, np.pi, B)
    x0 = np.stack([np.cos(th), np.sin(th), rng.uniform(-6, 6, B)], -1)
    acts = rng.uniform(-0.95, 0.95, (B, H, 1))
    g_r, g_n = rng.standard_normal((B, H)), rng.standard_normal((B, H, 3))
    g_o, g_a = rng.standard_normal((B, H, 3)), rng.standard_normal((B, H, 1))
    p = orc.PendulumParams()
    xt = torch.tensor(x0, requires_grad=True)
    at = torch.tensor(acts, requires_grad=True)
    x, obs, nxt, rew = xt, [], [], []
    for t in range(H):
        a = at[:, t, 0]
        obs.append(x)
        thx = torch.atan2(x[:, 1], x[:, 0])
        d = torch.remainder(thx - p.target_angle + np.pi, 2 * np.pi) - np.pi
        rew.append(-(p.angle_cost * d ** 2 + 0.1 * x[:, 2] ** 2) - p.control_cost * a ** 2)
        thdd = 3 * p.g / (2 * p.l) * torch.sin(thx) + 3.0 / (p.m * p.l ** 2) * torch.clamp(a, -1, 1) * p.max_torque
        nw = torch.clamp(x[:, 2] + thdd * p.dt, -p.max_speed, p.max_speed)
        nth = thx + nw * p.dt
        x = torch.stack([torch.cos(nth), torch.sin(nth), nw], -1)
        nxt.append(x)
    obs, nxt, rew = torch.stack(obs, 1), torch.stack(nxt, 1), torch.stack(rew, 1)
    loss = (rew * torch.tensor(g_r)).sum() + (nxt * torch.tensor(g_n)).sum() + (obs * torch.tensor(g_o)).sum() + \
        (at * torch.tensor(g_a)).sum()
    want_a, want_x0 = torch.autograd.grad(loss, [at, xt])
    got_a, got_x0 = orc.rollout_policy_vjp(obs.detach().numpy(), acts, g_r, g_n, g_o, g_a, dtype=np.float64)
    np.testing.assert_allclose(got_a, want_a.numpy(), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(got_x0, want_x0.numpy(), rtol=1e-9, atol=1e-10)
    # lambda_return and its transpose
    r, nv = torch.tensor(rng.standard_normal((B, H)), requires_grad=True), torch.tensor(rng.standard_normal((B, H)), requires_grad=True)
    inputs = r + 0.99 * nv * (1 - 0.95)
    agg, rets = nv[:, -1], []
    for t in range(H - 1, -1, -1):
        agg = inputs[:, t] + 0.99 * 0.95 * agg
        rets.append(agg)
    rets = torch.stack(rets[::-1], 1)
    np.testing.assert_allclose(orc.lambda_return(r.detach().numpy(), nv.detach().numpy(), 0.99, 0.95, dtype=np.float64),
                               rets.detach().numpy(), rtol=1e-12)
    wr, wnv = torch.autograd.grad((rets * torch.tensor(g_r)).sum(), [r, nv])
    gr, gnv = orc.lambda_return_vjp(g_r, 0.99, 0.95, dtype=np.float64)
    np.testing.assert_allclose(gr, wr.numpy(), rtol=1e-10)
    np.testing.assert_allclose(gnv, wnv.numpy(), rtol=1e-10)
    # lambda = 1: discounted Monte Carlo return; lambda = 0: one-step return (optimizer_utils.py:125-126)
    r1, nv1 = rng.standard_normal(H).astype(np.float32), rng.standard_normal(H).astype(np.float32)
    np.testing.assert_allclose(orc.lambda_return(r1, nv1, 0.9, 0.0), r1 + np.float32(0.9) * nv1, rtol=1e-6)
    mc = sum(0.9 ** k * r1[k] for k in range(H)) + 0.9 ** H * nv1[-1]
    np.testing.assert_allclose(orc.lambda_return(r1, nv1, 0.9, 1.0)[0], mc, rtol=1e-5)
