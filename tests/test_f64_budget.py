"""The float32 oracle inside the float64 error budgets (tests/f64_truth.py).

The GPU tests assert the SAME budgets for the kernels (tests/test_gpu_parity.py), so both float32 evaluations are
shown to lie in one stated ball around the float64 value of the reference's formulas -- the external bound the
kernel-vs-oracle comparisons lack on their own (JAX cannot run here; the reference pins no float).
"""
import numpy as np
import pytest

import f64_truth as ft
from oracle import jax_prng as ojr
from oracle import mbpo_oracle as orc


def _states(n, seed):
    rng = np.random.default_rng(seed)
    th, w = rng.uniform(-np.pi, np.pi, n), rng.uniform(-8, 8, n)
    return np.stack([np.cos(th), np.sin(th), w], -1).astype(np.float32)


def test_oracle_step_inside_budget(budget_report):
    x = _states(400_000, 0)
    u = np.random.default_rng(1).uniform(-1.5, 1.5, x.shape[0]).astype(np.float32)    # beyond the torque clip too
    xn, r = orc.pendulum_step(x, u)
    fx, fr = ft.step_errors(xn, r, x, u)
    budget_report("oracle32/system_step", state_frac=fx, reward_frac=fr, n=x.shape[0])
    assert fx <= 1.0 and fr <= 1.0, (fx, fr)
    # the states the rollouts actually visit: theta near +-pi (the wrap of the reward's floored mod) and saturated speed
    th = np.concatenate([np.pi - np.logspace(-7, -1, 500), -np.pi + np.logspace(-7, -1, 500), np.logspace(-8, -2, 500)])
    xe = np.stack([np.cos(th), np.sin(th), np.tile([8.0, -8.0, 0.0], 500)], -1).astype(np.float32)
    ue = np.tile([1.0, -1.0, 0.3], 500).astype(np.float32)
    xn, r = orc.pendulum_step(xe, ue)
    fx, fr = ft.step_errors(xn, r, xe, ue)
    budget_report("oracle32/system_step_edges", state_frac=fx, reward_frac=fr, n=xe.shape[0])
    assert fx <= 1.0 and fr <= 1.0, (fx, fr)


def test_budget_can_fail():
    """The budget is a real constraint: a step off by 2e-6 in the angle, or a reward off by 3e-5 relative, is outside."""
    x = _states(1000, 2)
    u = np.zeros(1000, np.float32)
    xn, r = orc.pendulum_step(x, u)
    th = np.arctan2(xn[:, 1], xn[:, 0]) + 2e-6
    bad = np.stack([np.cos(th), np.sin(th), xn[:, 2]], -1).astype(np.float32)
    fx, _ = ft.step_errors(bad, r, x, u)
    assert fx > 1.0
    _, fr = ft.step_errors(xn, r * np.float32(1 + 3e-5), x, u)
    assert fr > 1.0


@pytest.mark.parametrize("horizon", [7, 20, 30, 50])
def test_oracle_rollout_inside_budget(horizon, budget_report):
    """Open-loop returns: every float32 return lies within the first-order float64 amplification bound of its own
    rollout; the rows beyond the north star's rel 1e-5 are counted (they exist from H = 30 on: SURVEY section 7)."""
    R = 4000
    x0 = _states(R, 3)
    acts = np.clip(np.random.default_rng(4).normal(0, 0.5, (R, horizon)), -1, 1).astype(np.float32)
    r32 = orc.rollout_actions(x0, acts)
    r64, bound = ft.rollout_return_budget(x0, acts)
    err = np.abs(r32 - r64)
    frac = err / bound
    beyond = int((err > 1e-5 * np.abs(r64) + 1e-6).sum())
    budget_report("oracle32/rollout_return_H%d" % horizon, max_frac=float(frac.max()), median_frac=float(np.median(frac)),
                  rows=R, rows_beyond_rel_1e5=beyond, max_rel_err=float((err / np.abs(r64)).max()),
                  max_bound_rel=float((bound / np.abs(r64)).max()))
    assert frac.max() <= 1.0
    assert np.all(bound[err > 1e-5 * np.abs(r64) + 1e-6] > 1e-5 * np.abs(r64[err > 1e-5 * np.abs(r64) + 1e-6]))
    # per-observation forward bound
    _, obs32, _, _ = orc.rollout_actions(x0, acts, full=True)
    obs64, e = ft.rollout_state_budget(x0, acts)
    assert np.all(np.abs(obs32 - obs64)[:, 1:] <= e[:, 1:])


@pytest.mark.parametrize("horizon", [5, 8, 15, 20, 30, 50])
@pytest.mark.parametrize("exponent", [0.0, 2.0])
def test_oracle_noise_inside_budget(horizon, exponent, budget_report):
    keys = np.random.default_rng(horizon).integers(0, 2 ** 32, size=(2000, 2), dtype=np.uint64).astype(np.uint32)
    y, br, bi = orc.powerlaw_psd_gaussian_keys(exponent, horizon, keys, return_bits=True)
    z32, z64 = ojr.bits_to_normal(br), ft.normal_truth(br)
    fz = float((np.abs(z32 - z64) / (ft.NORMAL_REL * np.abs(z64) + ft.NORMAL_ABS)).max())
    truth = ft.powerlaw_truth(exponent, horizon, br, bi)
    fy = float(np.abs(y - truth).max() / ft.noise_budget(horizon))
    budget_report("oracle32/noise_H%d_exp%g" % (horizon, exponent), normal_frac=fz, noise_frac=fy)
    assert fz <= 1.0 and fy <= 1.0, (fz, fy)


def test_flip_explained_logic():
    v = np.array([0.0, 1.0, 2.0, 2.0000001, 3.0, 4.0])
    gap = np.full(6, 1e-6)
    assert ft.flip_explained(v, [3, 4, 5], [2, 4, 5], gap)            # boundary tie within the gap
    assert not ft.flip_explained(v, [1, 4, 5], [3, 4, 5], gap)        # 1.0 vs 2.0: not a rounding matter
    assert ft.flip_explained(v, [2, 3, 5], [3, 2, 5], gap)            # rank swap inside the gap
    assert not ft.flip_explained(v, [2, 5, 4], [2, 4, 5], gap)        # different tracked best
