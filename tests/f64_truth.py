"""Float64 truth leg for the float path (test infrastructure).

The reference pins no float (tests/test_icemopt.py:37-38 is a threshold) and JAX cannot run here, so a
kernel-vs-oracle comparison alone is float32 against float32.  This module evaluates the SAME formulas
(oracle/mbpo_oracle.py with ``dtype=float64``, which follow pendulum_dynamics.py:29-63, pendulum_reward.py:27-42,
optimizer_utils.py:28-58, general_utils.py:143-207) in float64 from the same float32 inputs and states an error
budget per operation.  A faithful float32 implementation -- the CUDA kernels, the NumPy oracle, XLA -- evaluates each
formula with a handful of roundings, so all of them lie in the same ball around the float64 value; the tests assert
that the kernels AND the float32 oracle are inside it.

Budgets are in units of u = 2**-24 (float32 unit roundoff):

  System.step, one step from a float32 state (|cos|,|sin| <= 1, |thdot| <= max_speed = 8):
    cos', sin'   10u                  atan2 (~1 ulp of pi = 4u) + th + thdot'*dt (2u at |th| <= 4) + cos/sin (1u),
                                      propagated with |d cos / d th| <= 1
    thdot'       12u                  one rounding of thdd*dt + thdot at magnitude <= 8 (8u) + the clip is exact,
                                      + c_g * sin(th) carried through dt (14.7 * 0.05 * 2u)
    reward       14u * |r| + 8u       diff**2 + 0.1*thdot**2 + 0.02*u**2, four roundings of terms <= |r|
                                      plus 2*|diff|*err(diff), err(diff) ~ 6u (atan2 and the +pi / -pi pair)

  Open-loop return over H steps (mean reward): first-order propagation of the per-step budgets through the float64
  adjoint of the rollout (lambda_t = d return / d x_t): see ``rollout_return_budget``.  This is what "the pendulum
  is chaotic" means quantitatively: near the upright equilibrium |lambda| grows like exp(t * dt / 0.26 s).
"""
from __future__ import annotations

import numpy as np

from oracle import mbpo_oracle as orc

U = 2.0 ** -24
STEP_CS_U = 10.0
STEP_W_U = 12.0
STEP_R_REL = 14.0 * U
STEP_R_ABS = 8.0 * U
F64 = np.float64


def step_truth(x32: np.ndarray, u32: np.ndarray, p=orc.PendulumParams()):
    """float64 System.step of float32 inputs -> (x_next64 [n,3], reward64 [n])."""
    return orc.pendulum_step(np.asarray(x32, F64), np.asarray(u32, F64), p, dtype=F64)


def step_budget(x_next64: np.ndarray, reward64: np.ndarray):
    """Per-element error budgets of one float32 step around the float64 value."""
    bx = np.empty_like(x_next64)
    bx[..., 0] = STEP_CS_U * U
    bx[..., 1] = STEP_CS_U * U
    bx[..., 2] = STEP_W_U * U
    br = STEP_R_REL * np.abs(reward64) + STEP_R_ABS
    return bx, br


def step_errors(x_next: np.ndarray, reward: np.ndarray, x32: np.ndarray, u32: np.ndarray, p=orc.PendulumParams()):
    """Errors of a float32 step result against the float64 truth, as fractions of the budget (<= 1 passes).
    Returns (worst state fraction, worst reward fraction)."""
    xn64, r64 = step_truth(x32, u32, p)
    bx, br = step_budget(xn64, r64)
    ex = np.abs(np.asarray(x_next, F64) - xn64) / bx
    er = np.abs(np.asarray(reward, F64) - r64) / br
    return float(ex.max()) if ex.size else 0.0, float(er.max()) if er.size else 0.0


def rollout_truth(x0_32: np.ndarray, actions32: np.ndarray, p=orc.PendulumParams()):
    """float64 open-loop rollout: x0 [R,3] (or [3]), actions [R,H] -> (returns [R], obs [R,H,3], rewards [R,H])."""
    ret, obs, rew, _ = orc.rollout_actions(np.asarray(x0_32, F64), np.asarray(actions32, F64), p, dtype=F64, full=True)
    return ret, obs, rew


def rollout_return_budget(x0_32: np.ndarray, actions32: np.ndarray, p=orc.PendulumParams()):
    """float64 returns of open-loop rollouts and, per rollout, the first-order bound on what a faithful float32
    evaluation may differ by:

        |R32 - R64| <= (1/H) sum_t b_r(t)  +  sum_t |lambda_{t+1}| . b_x(t)  +  H * u * max_t |partial sum| / H

    with b_x, b_r the per-step budgets above, lambda_t = d R / d x_t from the float64 reverse pass
    (oracle.pendulum_step_vjp, itself pinned against central differences), and the last term the H float32
    additions of the horizon mean.  Returns (R64 [R], bound [R])."""
    a = np.asarray(actions32, F64)
    r, h = a.shape
    ret, obs, rew = rollout_truth(x0_32, actions32, p)
    lam = np.zeros((r, 3), F64)
    bound = np.zeros(r, F64)
    g_r = np.full(r, 1.0 / h, F64)
    for t in range(h - 1, -1, -1):
        # budget of the step t -> t+1 result enters through lambda_{t+1} (zero after the last step: the final
        # next state is not part of the return)
        if t + 1 < h:
            nxt64 = obs[:, t + 1]
            bx, _ = step_budget(nxt64, rew[:, t])
            bound += (np.abs(lam) * bx).sum(-1)
        _, br = step_budget(obs[:, t], rew[:, t])
        bound += br / h
        lam, _ = orc.pendulum_step_vjp(obs[:, t], a[:, t], lam, g_r, p, dtype=F64)
    partial = np.abs(np.cumsum(rew, axis=1)).max(axis=1)
    bound += (h + 1) * U * partial / h
    return ret, bound


def rollout_state_budget(x0_32: np.ndarray, actions32: np.ndarray, p=orc.PendulumParams()):
    """Forward companion of rollout_return_budget: a first-order bound on the error of every observation of a
    float32 open-loop rollout, e_0 = 0, e_{t+1} = |J_t| e_t + b_x with J_t the float64 Jacobian of System.step
    (rows from oracle.pendulum_step_vjp with unit cotangents).  Returns (obs64 [R,H,3], e [R,H,3])."""
    a = np.asarray(actions32, F64)
    r, h = a.shape
    _, obs, rew = rollout_truth(x0_32, actions32, p)
    e = np.zeros((r, h, 3), F64)
    zero_r = np.zeros(r, F64)
    for t in range(h - 1):
        nxt = np.zeros((r, 3), F64)
        for k in range(3):
            unit = np.zeros((r, 3), F64)
            unit[:, k] = 1.0
            row, _ = orc.pendulum_step_vjp(obs[:, t], a[:, t], unit, zero_r, p, dtype=F64)   # d x'_k / d x
            nxt[:, k] = (np.abs(row) * e[:, t]).sum(-1)
        bx, _ = step_budget(obs[:, t + 1], rew[:, t])
        e[:, t + 1] = nxt + bx
    return obs, e


def flip_explained(values: np.ndarray, idx_a: np.ndarray, idx_b: np.ndarray, gap: np.ndarray) -> bool:
    """Two elite selections from (nearly) the same values may differ only where the values cannot be told
    apart: every candidate one side selected and the other did not must be within ``gap`` of the selection
    boundary, and any two elites ranked differently must be within ``gap`` of each other.  ``gap`` [M]: the error
    budget of each value."""
    values = np.asarray(values, F64)
    a, b = [int(i) for i in idx_a], [int(i) for i in idx_b]
    sa, sb = set(a), set(b)
    only = sorted(sa ^ sb)
    if only:
        kth = min(values[a].min(), values[b].min())          # the boundary both selections end at
        for i in only:
            if abs(values[i] - kth) > gap[i] + gap[only].max():
                return False
    common_a = [i for i in a if i in sb]
    common_b = [i for i in b if i in sa]
    for i, j in zip(common_a, common_b):                      # same rank, different candidate
        if i != j and abs(values[i] - values[j]) > gap[i] + gap[j]:
            return False
    if a[-1] != b[-1] and abs(values[a[-1]] - values[b[-1]]) > gap[a[-1]] + gap[b[-1]]:
        return False                                          # the tracked best (:217-226)
    return True


# ---- colored noise -----------------------------------------------------------------------------------------------
def normal_truth(bits: np.ndarray) -> np.ndarray:
    """jax.random.normal of 32-bit words with every operation in float64 EXCEPT the two float32 roundings the
    formula's own definition fixes: u = f * 2 + nextafter(-1, 0) (one float32 rounding; f * 2 is exact) and
    t = float32(u * u).  XLA's ErfInv32 evaluates -log1p(-t) from that rounded t: for |u| -> 1 the rounding of
    u * u is a several-percent perturbation of 1 - u * u, which every float32 implementation shares -- a float64
    "truth" without it would describe a different function."""
    from oracle import jax_prng as jr
    u32 = jr.bits_to_uniform(bits, jr._NORMAL_LO, np.float32(1.0))
    t = (u32 * u32).astype(np.float32).astype(F64)
    x = u32.astype(F64)
    w = -np.log1p(-t)
    small = w < 5.0
    ww = np.where(small, w - 2.5, np.sqrt(np.maximum(w, 0.0)) - 3.0)
    cs, cl = jr._ERFINV_SMALL.astype(F64), jr._ERFINV_LARGE.astype(F64)
    p = np.where(small, cs[0], cl[0])
    for i in range(1, 9):
        p = np.where(small, cs[i], cl[i]) + p * ww
    return F64(np.float32(np.sqrt(2))) * (p * x)


NORMAL_REL = 8 * U          # Horner of 9 float32 terms + log1p + two multiplications
NORMAL_ABS = 1e-8


def powerlaw_truth(exponent: float, size: int, bits_r: np.ndarray, bits_i: np.ndarray) -> np.ndarray:
    """powerlaw_psd_gaussian (general_utils.py:143-207) in float64 from the PRNG words: tables, normals
    (normal_truth), DC / Nyquist fix, irfft, division by sigma.  bits_* [M, F] -> y [M, size]."""
    s_scale, sigma = orc.powerlaw_tables(exponent, size, dtype=F64)
    sr = normal_truth(bits_r) * s_scale
    si = normal_truth(bits_i) * s_scale
    if size % 2 == 0:
        si[:, -1] = 0
        sr[:, -1] = sr[:, -1] * np.sqrt(2.0)
    si[:, 0] = 0
    sr[:, 0] = sr[:, 0] * np.sqrt(2.0)
    return np.fft.irfft(sr + 1j * si, n=size, axis=-1) / sigma


def noise_budget(size: int, exponent: float = 0.0) -> float:
    """Absolute budget of one unit-variance noise sample: F = size//2+1 products of a float32 normal (8u
    relative), a float32 table entry and a float32 twiddle, accumulated in float32 (one rounding per term of a
    partial sum of magnitude <~ 4): 6u per bin, plus 4 bins' worth for the normalisation and the
    final combination of the partial sums (what is left when H = 2 has two bins).  Measured: the float32 oracle and the kernels use <= 0.6 of it
    (profiles/r02_f64_budget_report_*.json)."""
    return (size // 2 + 5) * 6.0 * U
