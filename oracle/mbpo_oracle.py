"""ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the mbpo
iCEM planning hot path, following the reference line by line.

Reference files restated here (paths relative to /root/reference):
  mbpo/utils/general_utils.py:81-208                       powerlaw_psd_gaussian
  mbpo/systems/dynamics/pendulum_dynamics.py:12-63         PendulumDynamics.next_state / ode
  mbpo/systems/rewards/pendulum_reward.py:12-42            PendulumReward.__call__
  mbpo/systems/pendulum_system.py:18-46                    PendulumSystem.step / reset
  mbpo/utils/optimizer_utils.py:11-59                      rollout_actions
  mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:121-257   iCemTO.init/optimize/act
  mbpo/systems/brax_wrapper.py:40-50, mbpo/optimizers/policy_optimizers/brax_utils/
  training.py:71-137, .../sac/acting.py:35-55              vmapped env step + wrappers
  mbpo/utils/network_utils.py:5-17                         MLP template (learned dynamics)
  mbpo/optimizers/policy_optimizers/sac/sac_networks.py:58-73, sac/parametric_distribution.py:66-125,
  ppo/ppo_network.py:59-84, sac/sac.py:283-292             policy in the env loop (NormalTanh sample, log_prob)
  mbpo/utils/optimizer_utils.py:62-131                     rollout_policy, lambda_return
  mbpo/optimizers/policy_optimizers/bptt_optimizer.py:107-142,306-376   BPTT actor / act / the differentiated rollout
    (the cotangent pass is derived by hand here and pinned against finite differences and torch autograd,
    tests/test_oracle_bptt.py; distrax's Normal.log_prob / Tanh.forward_log_det_jacobian are restated from the
    published distrax 0.1.x formulas -- distrax is a third-party dependency absent from /root/reference)

PARITY PINNING: the reference is pure JAX, JAX cannot be installed here, and the
reference's tests hold no golden vectors (tests/test_icemopt.py:37-38 is the threshold
sum(rewards) >= -400).  This restatement is therefore pinned by (i) the JAX PRNG
known-answer vectors (oracle/jax_prng.py), (ii) the reference's behavioural threshold,
(iii) an independent plain-C twin (oracle/c/mbpo_oracle.c).  Float parity against a
real JAX/XLA run is UNPINNED.

Summation orders that XLA leaves unspecified (mean over the horizon, mean over elites)
are fixed here as plain left-to-right float32 sums.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
import this module.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Optional

import numpy as np

from . import jax_prng as jr

F32 = np.float32
U32 = np.uint32


# ----------------------------------------------------------------------------------
# iCemParams  (icem_optimizer.py:25-50)
# ----------------------------------------------------------------------------------
@dataclass(frozen=True)
class ICemParams:
    num_particles: int = 10
    num_samples: int = 500
    num_elites: int = 50
    init_std: float = 0.5
    alpha: float = 0.0
    num_steps: int = 5
    exponent: float = 0.0
    elite_set_fraction: float = 0.3
    u_min: float = -1.0
    u_max: float = 1.0
    warm_start: bool = True
    lambda_constraint: float = 1e4

    @property
    def num_prev_elites(self) -> int:  # icem_optimizer.py:170
        return max(int(self.elite_set_fraction * self.num_elites), 1)


# ----------------------------------------------------------------------------------
# Pendulum parameters  (pendulum_dynamics.py:12-19, pendulum_reward.py:12-16)
# ----------------------------------------------------------------------------------
@dataclass(frozen=True)
class PendulumParams:
    max_speed: float = 8.0
    max_torque: float = 2.0
    dt: float = 0.05
    g: float = 9.81
    m: float = 1.0
    l: float = 1.0
    control_cost: float = 0.02
    angle_cost: float = 1.0
    target_angle: float = 0.0

    def packed(self) -> np.ndarray:
        return np.array([self.max_speed, self.max_torque, self.dt, self.g, self.m, self.l,
                         self.control_cost, self.angle_cost, self.target_angle], dtype=F32)


# ----------------------------------------------------------------------------------
# powerlaw_psd_gaussian  (general_utils.py:81-208)
# ----------------------------------------------------------------------------------
def powerlaw_tables(exponent: float, samples: int, dtype=F32):
    """Static (trace-time) part: s_scale[F] and sigma  (general_utils.py:143-178)."""
    dt = dtype
    nfreq = samples // 2 + 1
    f = np.arange(nfreq, dtype=dt) / dt(samples)                     # rfftfreq, :143
    fmin = max(0.0, 1.0 / samples)                                    # :146-147
    s_scale = f.copy()
    ix = int(np.sum(s_scale < dt(fmin)))                              # :153
    if ix and ix < len(s_scale):                                      # :166-172
        s_scale[:ix] = s_scale[ix]
    s_scale = np.power(s_scale, dt(-exponent / 2.0)).astype(dt)       # :173
    w = s_scale[1:].copy()                                            # :176
    w[-1] = w[-1] * dt((1 + (samples % 2)) / 2.0)                     # :177
    sigma = dt(2) * np.sqrt(np.sum(w ** 2, dtype=dt)) / dt(samples)   # :178
    return s_scale.astype(dt), dt(sigma)


def irfft_direct(sr, si, n, dtype=F32):
    """Direct real inverse DFT of the half spectrum (jnp.fft.irfft semantics: the
    imaginary parts of the DC and, for even n, the Nyquist bins are ignored)."""
    nfreq = n // 2 + 1
    t = np.arange(n)
    k = np.arange(nfreq)
    ang = 2.0 * np.pi * ((np.outer(k, t)) % n) / n
    c = np.cos(ang)
    s = np.sin(ang)
    wgt = np.full(nfreq, 2.0)
    wgt[0] = 1.0
    if n % 2 == 0:
        wgt[-1] = 1.0
    cr = (wgt[:, None] * c / n).astype(dtype)
    ci = (-wgt[:, None] * s / n).astype(dtype)
    ci[0] = 0
    if n % 2 == 0:
        ci[-1] = 0
    return (sr.astype(dtype) @ cr + si.astype(dtype) @ ci).astype(dtype)


def powerlaw_psd_gaussian_keys(exponent: float, size: int, keys: np.ndarray,
                               partitionable: bool = False, dtype=F32,
                               return_bits: bool = False):
    """vmap(powerlaw_psd_gaussian) over keys uint32[M,2] -> float[M, size].

    Per key (general_utils.py:189-207): key_sr, key_si, _ = split(rng, 3);
    sr = normal(key_sr,(F,))*s_scale; si likewise; Nyquist/DC fix; irfft(n=size)/sigma.
    """
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    m = keys.shape[0]
    nfreq = size // 2 + 1
    s_scale, sigma = powerlaw_tables(exponent, size, dtype)
    sr = np.empty((m, nfreq), dtype=F32)
    si = np.empty((m, nfreq), dtype=F32)
    bits_r = np.empty((m, nfreq), dtype=U32)
    bits_i = np.empty((m, nfreq), dtype=U32)
    sub = split_keys(keys, 3, partitionable)                          # :189
    for out, bits_out, kk in ((sr, bits_r, sub[:, 0]), (si, bits_i, sub[:, 1])):
        b = random_bits_keys(kk, nfreq, partitionable)
        bits_out[:] = b
        out[:] = jr.bits_to_normal(b)                                 # :190-191
    sr = sr.astype(dtype) * s_scale
    si = si.astype(dtype) * s_scale
    if size % 2 == 0:                                                 # :195-197
        si[:, -1] = 0
        sr[:, -1] = sr[:, -1] * dtype(np.sqrt(2))
    si[:, 0] = 0                                                      # :200-201
    sr[:, 0] = sr[:, 0] * dtype(np.sqrt(2))
    y = np.fft.irfft(sr + 1j * si, n=size, axis=-1).astype(dtype) / sigma   # :207
    if return_bits:
        return y.astype(dtype), bits_r, bits_i
    return y.astype(dtype)


def powerlaw_psd_gaussian(exponent: float, size: int, rng, partitionable: bool = False, dtype=F32):
    return powerlaw_psd_gaussian_keys(exponent, size, np.asarray(rng).reshape(1, 2),
                                      partitionable, dtype)[0]


# -- vmapped PRNG helpers ----------------------------------------------------------
def split_keys(keys: np.ndarray, num: int, partitionable: bool = False) -> np.ndarray:
    """vmap(lambda k: split(k, num))(keys): uint32[M,2] -> uint32[M,num,2]."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    k0 = keys[:, 0:1]
    k1 = keys[:, 1:2]
    if partitionable:
        y0, y1 = jr.threefry2x32(k0, k1, np.zeros((1, num), dtype=U32), np.arange(num, dtype=U32)[None])
        return np.stack([y0, y1], axis=-1)
    c = np.arange(2 * num, dtype=U32)
    y0, y1 = jr.threefry2x32(k0, k1, c[None, :num], c[None, num:])
    return np.concatenate([y0, y1], axis=1).reshape(-1, num, 2)


def random_bits_keys(keys: np.ndarray, n: int, partitionable: bool = False) -> np.ndarray:
    """vmap(lambda k: random_bits(k,(n,)))(keys): uint32[M,2] -> uint32[M,n]."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    k0 = keys[:, 0:1]
    k1 = keys[:, 1:2]
    if partitionable:
        y0, y1 = jr.threefry2x32(k0, k1, np.zeros((1, n), dtype=U32), np.arange(n, dtype=U32)[None])
        return y0 ^ y1
    npad = n + (n % 2)
    c = np.arange(npad, dtype=U32)
    if n % 2:
        c[-1] = 0
    h = npad // 2
    y0, y1 = jr.threefry2x32(k0, k1, c[None, :h], c[None, h:])
    return np.concatenate([y0, y1], axis=1)[:, :n]


# ----------------------------------------------------------------------------------
# Pendulum System.step  (pendulum_system.py:18-39; batched over leading axis)
# ----------------------------------------------------------------------------------
def pendulum_step(x: np.ndarray, u: np.ndarray, p: PendulumParams = PendulumParams(), dtype=F32):
    """x[...,3], u[...] (scalar action per row) -> (x_next[...,3], reward[...])."""
    d = dtype
    x = np.asarray(x, dtype=d)
    u = np.asarray(u, dtype=d)
    max_speed, max_torque, dt, g, m, l = (d(p.max_speed), d(p.max_torque), d(p.dt), d(p.g), d(p.m), d(p.l))
    th = np.arctan2(x[..., 1], x[..., 0]).astype(d)                   # dynamics :35
    thdot = x[..., 2]
    uu = (np.clip(u, d(-1), d(1)) * max_torque).astype(d)             # :59
    c_g = d(d(d(3) * g) / d(d(2) * l))                                # 3*g/(2*l)
    c_u = d(d(3.0) / d(m * d(l * l)))                                 # 3.0/(m*l**2)
    thdd = (c_g * np.sin(th).astype(d) + c_u * uu).astype(d)          # :60
    nthd_ode = np.clip((thdot + thdd * dt).astype(d), -max_speed, max_speed)   # :61-62
    newth = (th + nthd_ode * dt).astype(d)                            # :40
    nthd = np.clip((thdot + thdd * dt).astype(d), -max_speed, max_speed)       # :41-42
    x_next = np.stack([np.cos(newth), np.sin(newth), nthd], axis=-1).astype(d)  # :43
    # reward on the CURRENT state and the raw action  (pendulum_reward.py:32-40)
    pi = d(np.pi)
    two_pi = d(2 * np.pi)
    diff = (th - d(p.target_angle)).astype(d)
    diff = (np.remainder((diff + pi).astype(d), two_pi).astype(d) - pi).astype(d)   # :35 floored mod
    reward = (-(d(p.angle_cost) * (diff * diff) + d(0.1) * (thdot * thdot))
              - d(p.control_cost) * (u * u)).astype(d)                # :38-39
    return x_next, reward


def pendulum_reset(dtype=F32):
    """PendulumSystem.reset (pendulum_system.py:41-46)."""
    return np.array([-1.0, 0.0, 0.0], dtype=dtype), dtype(0.0)


# ----------------------------------------------------------------------------------
# Learned MLP-ensemble System (config 4; template network_utils.py:5-17, swish)
# ----------------------------------------------------------------------------------
@dataclass
class MlpEnsembleParams:
    """E members of [X+A -> hidden.. -> X] with swish; predicts delta-x.  Weights are
    stored [E, in, out] (flax Dense kernel layout), biases [E, out]."""
    weights: list
    biases: list
    reward: PendulumParams = field(default_factory=PendulumParams)

    @property
    def num_members(self) -> int:
        return self.weights[0].shape[0]


def make_mlp_ensemble(seed: int = 3, members: int = 5, x_dim: int = 3, u_dim: int = 1,
                      hidden=(256, 256, 256), scale: float = 1.0) -> MlpEnsembleParams:
    rng = np.random.default_rng(seed)
    dims = (x_dim + u_dim,) + tuple(hidden) + (x_dim,)
    ws, bs = [], []
    for i in range(len(dims) - 1):
        ws.append((rng.standard_normal((members, dims[i], dims[i + 1])) * scale / np.sqrt(dims[i])).astype(F32))
        bs.append((0.01 * rng.standard_normal((members, dims[i + 1]))).astype(F32))
    # keep the predicted delta small so rollouts stay bounded
    ws[-1] = (ws[-1] * F32(0.1)).astype(F32)
    return MlpEnsembleParams(ws, bs)


def _bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bfloat16 -> float32."""
    a = np.ascontiguousarray(a, dtype=F32)
    b = a.view(U32)
    rounded = ((b + U32(0x7FFF) + ((b >> U32(16)) & U32(1))) & U32(0xFFFF0000)).astype(U32)
    return rounded.view(F32)


def swish(x):
    return x / (1.0 + np.exp(-x))


def mlp_member_forward(params: MlpEnsembleParams, member: np.ndarray, inp: np.ndarray,
                       bf16: bool = False) -> np.ndarray:
    """inp[M, X+A], member[M] (which ensemble member serves each row) -> delta[M, X].

    bf16=True applies the same operand rounding as the tensor-core path: activations and
    hidden-layer weights rounded to bfloat16, products accumulated in float32; the first
    layer (K = X+A) stays float32 and the last layer (N = X) takes bfloat16 activations against
    float32 weights.
    """
    h = np.asarray(inp, dtype=F32)
    nl = len(params.weights)
    out = np.empty((h.shape[0], params.weights[-1].shape[-1]), dtype=F32)
    for e in range(params.num_members):
        rows = np.nonzero(member == e)[0]
        if rows.size == 0:
            continue
        a = h[rows]
        for li in range(nl):
            w = params.weights[li][e]
            b = params.biases[li][e]
            hidden_gemm = 0 < li < nl - 1
            if bf16 and hidden_gemm:
                a = (_bf16_round(a).astype(np.float64) @ _bf16_round(w).astype(np.float64)).astype(F32) + b
            elif bf16 and li == nl - 1:
                # output layer: bf16 activations against the float32 weights (the fused kernel feeds the
                # tensor core w = bf16(w) + bf16(w - bf16(w)), i.e. w to 2^-17)
                a = (_bf16_round(a).astype(np.float64) @ w.astype(np.float64)).astype(F32) + b
            else:
                a = (a.astype(np.float64) @ w.astype(np.float64)).astype(F32) + b
            if li < nl - 1:
                a = swish(a.astype(F32)).astype(F32)
        out[rows] = a
    return out


def mlp_ensemble_step(x, u, member, params: MlpEnsembleParams, bf16: bool = False):
    """System.step for the learned ensemble: x_next = x + MLP_member([x,u]); the reward is
    the pendulum reward (pendulum_reward.py:27-42) on the current state."""
    x = np.asarray(x, dtype=F32)
    u = np.asarray(u, dtype=F32)
    inp = np.concatenate([x, u[..., None]], axis=-1).astype(F32)
    dx = mlp_member_forward(params, member, inp, bf16)
    x_next = (x + dx).astype(F32)
    _, reward = pendulum_step(x, u, params.reward)
    return x_next, reward


def ensemble_rollout_returns(x0, actions, params: MlpEnsembleParams, bf16: bool = True, use_max: bool = False):
    """vmap(objective) through the learned ensemble (icem_optimizer.py:144-160): x0[B,3],
    actions[B,M,H] -> [B,M].  Particle p is rolled through member p; the objective is the
    mean (or max) over members of the horizon-mean reward (left-to-right float32 sums)."""
    x0 = np.asarray(x0, dtype=F32)
    actions = np.asarray(actions, dtype=F32)
    b, m, h = actions.shape
    rows = b * m
    acts = actions.reshape(rows, h)
    per_member = []
    for e in range(params.num_members):
        x = np.repeat(x0, m, axis=0).copy()
        member = np.full(rows, e, dtype=np.int32)
        acc = np.zeros(rows, dtype=F32)
        for t in range(h):
            xn, r = mlp_ensemble_step(x, acts[:, t], member, params, bf16)
            acc = (acc + r).astype(F32)
            x = xn
        per_member.append((acc / F32(h)).astype(F32))
    if use_max:
        out = per_member[0]
        for r in per_member[1:]:
            out = np.maximum(out, r)
    else:
        out = np.zeros(rows, dtype=F32)
        for r in per_member:
            out = (out + r).astype(F32)
        out = (out / F32(params.num_members)).astype(F32)
    return out.reshape(b, m)


# ----------------------------------------------------------------------------------
# rollout_actions  (optimizer_utils.py:11-59), batched over rows
# ----------------------------------------------------------------------------------
def rollout_actions(x0: np.ndarray, actions: np.ndarray, p: PendulumParams = PendulumParams(),
                    dtype=F32, full: bool = False):
    """x0[R,3] (or [3]), actions[R,H] -> mean-over-horizon reward[R]  (icem :160 inner mean).

    full=True also returns the brax Transition fields (observation, reward,
    next_observation) as arrays [R,H,3], [R,H], [R,H,3]  (optimizer_utils.py:47-58).
    """
    actions = np.asarray(actions, dtype=dtype)
    r, h = actions.shape
    x = np.broadcast_to(np.asarray(x0, dtype=dtype), (r, 3)).copy()
    acc = np.zeros(r, dtype=dtype)
    if full:
        obs = np.empty((r, h, 3), dtype=dtype)
        nxt = np.empty((r, h, 3), dtype=dtype)
        rew = np.empty((r, h), dtype=dtype)
    for t in range(h):                                                # lax.scan :46
        xn, rw = pendulum_step(x, actions[:, t], p, dtype)
        if full:
            obs[:, t] = x
            nxt[:, t] = xn
            rew[:, t] = rw
        acc = (acc + rw).astype(dtype)                                # left-to-right f32 sum
        x = xn
    ret = (acc / dtype(h)).astype(dtype)                              # jnp.mean over horizon
    if full:
        return ret, obs, rew, nxt
    return ret


# ----------------------------------------------------------------------------------
# iCEM  (icem_optimizer.py:121-257)
# ----------------------------------------------------------------------------------
@dataclass
class ICemState:
    """iCemOptimizerState (icem_optimizer.py:62-69) minus the opaque replay buffer."""
    key: np.ndarray
    best_sequence: np.ndarray
    best_reward: np.ndarray

    @property
    def action(self):
        return self.best_sequence[0]

    def replace(self, **kw):
        return replace(self, **kw)


def icem_init(key, horizon: int, action_dim: int = 1, partitionable: bool = False) -> ICemState:
    """iCemTO.init (icem_optimizer.py:121-132): init_key, dummy_buffer_key, key = split(key, 3)."""
    ks = jr.split(key, 3, partitionable)
    return ICemState(key=ks[2], best_sequence=np.zeros((horizon, action_dim), dtype=F32),
                     best_reward=F32(0.0))


def total_order_key(v: np.ndarray) -> np.ndarray:
    """Monotone uint32 image of float32 under JAX's sort order (jax/_src/lax/lax.py
    _float_to_int_for_sort): IEEE total order with -0.0 == +0.0 and all NaNs equal, last."""
    v = np.asarray(v, dtype=F32)
    b = v.view(U32).copy()
    b[v == 0] = U32(0)
    b[np.isnan(v)] = U32(0x7FC00000)
    neg = (b >> U32(31)).astype(bool)
    return np.where(neg, ~b, b | U32(0x80000000)).astype(U32)


def stable_argsort(values: np.ndarray) -> np.ndarray:
    """jnp argsort: stable ascending under the total order (icem_optimizer.py:199)."""
    return np.argsort(total_order_key(values), kind="stable")


def icem_sample_actions(carry_key, mean, std, params: ICemParams, horizon: int, action_dim: int = 1,
                        partitionable: bool = False, dtype=F32, return_bits: bool = False):
    """One iteration of key plumbing + colored sampling  (icem_optimizer.py:174-192).

    Returns (next_key, actions[N+Np, H, A], particle_keys[N+Np, 2]) (+ bit streams)."""
    n = params.num_samples
    npe = params.num_prev_elites
    ks = jr.split(carry_key, 2, partitionable)                        # :174
    sampling_rng, particles_rng = ks[0], ks[1]
    srs = jr.split(sampling_rng, n + 1, partitionable)                # :175
    next_key, sample_keys = srs[0], srs[1:]                           # :176
    particle_keys = jr.split(particles_rng, n + npe, partitionable)   # :177
    dim_keys = split_keys(sample_keys, action_dim, partitionable)     # :180  [N, A, 2]
    res = powerlaw_psd_gaussian_keys(params.exponent, horizon, dim_keys.reshape(-1, 2),
                                     partitionable, dtype, return_bits=return_bits)
    noise = res[0] if return_bits else res
    colored = noise.reshape(n, action_dim, horizon).transpose(0, 2, 1)   # out_axes=1, :185-187
    mean = np.asarray(mean, dtype=dtype)
    std = np.asarray(std, dtype=dtype)
    acts = (mean[None] + colored * std[None]).astype(dtype)           # :190
    acts = np.clip(acts, np.asarray(params.u_min, dtype=dtype), np.asarray(params.u_max, dtype=dtype))   # :191
    prev_elites = np.zeros((npe, horizon, action_dim), dtype=dtype)   # closure zeros, :192,:245
    acts = np.concatenate([acts, prev_elites], axis=0)
    if return_bits:
        return next_key, acts, particle_keys, res[1], res[2]
    return next_key, acts, particle_keys


def particle_mean(value_per_particle: np.ndarray, dtype=F32) -> np.ndarray:
    """jnp.mean over the particle axis (last), left-to-right."""
    p = value_per_particle.shape[-1]
    acc = np.zeros(value_per_particle.shape[:-1], dtype=dtype)
    for i in range(p):
        acc = (acc + value_per_particle[..., i]).astype(dtype)
    return (acc / dtype(p)).astype(dtype)


def icem_objective(x0, acts, params: ICemParams, sys_params: PendulumParams, use_optimism=False, dtype=F32,
                   cost_fn=None, use_pessimism=False):
    """vmap(objective) (icem_optimizer.py:144-166,195) for the deterministic pendulum: every
    particle sees the same trajectory, so mean/max over particles is over P identical values.
    cost_fn(states [M,H,X], actions [M,H,A]) -> [M] is the (already vmapped) AbstractCost (:161-166)."""
    def summarize(v, use_max):
        if use_max or params.num_particles == 1:
            return v
        return particle_mean(np.repeat(v[:, None], params.num_particles, axis=1), dtype)
    if cost_fn is None:
        ret = rollout_actions(np.asarray(x0, dtype=dtype), acts[:, :, 0], sys_params, dtype)
        return summarize(ret, use_optimism)
    ret, obs, _, _ = rollout_actions(np.asarray(x0, dtype=dtype), acts[:, :, 0], sys_params, dtype, full=True)
    cost = np.asarray(cost_fn(obs, acts), dtype=dtype)
    cost = summarize(cost, use_pessimism)
    pen = (dtype(params.lambda_constraint) * np.maximum(cost, dtype(0))).astype(dtype)         # :166
    return (summarize(ret, use_optimism) - pen).astype(dtype)


def icem_refit(acts, values, mean, std, best_value, best_seq, params: ICemParams, dtype=F32):
    """Elite select + refit + best tracking  (icem_optimizer.py:199-226)."""
    k = params.num_elites
    idx = stable_argsort(values)[-k:]                                 # :199
    elites = acts[idx]                                                # :202
    elite_values = values[idx]                                        # :203
    acc = np.zeros(elites.shape[1:], dtype=dtype)
    for e in range(k):
        acc = (acc + elites[e]).astype(dtype)
    elite_mean = (acc / dtype(k)).astype(dtype)                       # :206
    acc = np.zeros(elites.shape[1:], dtype=dtype)
    for e in range(k):
        dlt = (elites[e] - elite_mean).astype(dtype)
        acc = (acc + dlt * dlt).astype(dtype)
    elite_var = (acc / dtype(k)).astype(dtype)                        # :207 (ddof=0, two-pass)
    a = dtype(params.alpha)
    one_m = dtype(1 - params.alpha)
    new_mean = (mean * a + one_m * elite_mean).astype(dtype)          # :210
    var = ((std * std) * a + one_m * elite_var).astype(dtype)         # :211
    new_std = np.sqrt(var).astype(dtype)                              # :214
    if best_value <= elite_values[-1]:                                # :217-226
        best_value, best_seq = elite_values[-1], elites[-1]
    return new_mean, new_std, best_value, best_seq, idx


def icem_optimize(x0, state: ICemState, params: ICemParams, horizon: int, action_dim: int = 1,
                  sys_params: PendulumParams = PendulumParams(), use_optimism: bool = False,
                  partitionable: bool = False, dtype=F32, trace: Optional[list] = None,
                  cost_fn=None, use_pessimism: bool = False) -> ICemState:
    """iCemTO.optimize (icem_optimizer.py:134-252) for one problem."""
    mean = np.zeros((horizon, action_dim), dtype=dtype)
    if params.warm_start:                                             # :239-241
        mean[:-1] = state.best_sequence[1:]
        mean[-1] = state.best_sequence[-1]
    std = np.full((horizon, action_dim), params.init_std, dtype=dtype)   # :243
    best_seq = mean.copy()                                            # :244
    best_value = dtype(-np.inf)                                       # :235
    ks = jr.split(state.key, 2, partitionable)                        # :246
    carry_key, new_state_key = ks[0], ks[1]
    for it in range(params.num_steps):                                # lax.scan :250
        in_key, in_mean, in_std = carry_key, mean, std
        carry_key, acts, _ = icem_sample_actions(carry_key, mean, std, params, horizon, action_dim,
                                                 partitionable, dtype)
        values = icem_objective(x0, acts, params, sys_params, use_optimism, dtype, cost_fn, use_pessimism)   # :195
        mean, std, best_value, best_seq, idx = icem_refit(acts, values, mean, std, best_value,
                                                          best_seq, params, dtype)
        if trace is not None:
            trace.append(dict(key=in_key, mean_in=in_mean, std_in=in_std, actions=acts, values=values,
                              elite_idx=idx, mean=mean, std=std, best_value=best_value,
                              best_seq=best_seq))
    return ICemState(key=new_state_key, best_sequence=np.asarray(best_seq, dtype=F32),
                     best_reward=F32(best_value))                     # :251


def icem_act(x0, state: ICemState, params: ICemParams, horizon: int, **kw):
    """iCemTO.act (icem_optimizer.py:254-257)."""
    new_state = icem_optimize(x0, state, params, horizon, **kw)
    return new_state.action, new_state


def closed_loop_mpc(num_steps: int = 200, horizon: int = 20, params: ICemParams = ICemParams(),
                    seed: int = 0, partitionable: bool = False, collapse_particles: bool = True):
    """tests/test_icemopt.py:6-32 -- returns (states[T,3], rewards[T], actions[T])."""
    key = jr.PRNGKey(seed)
    ks = jr.split(key, 3, partitionable)                              # optimizer_key, init_key, key
    init_key = ks[1]
    x, _ = pendulum_reset()
    st = icem_init(init_key, horizon, 1, partitionable)
    p = replace(params, num_particles=1) if collapse_particles else params
    xs, rs, us = [], [], []
    for _ in range(num_steps):
        a, st = icem_act(x, st, p, horizon, partitionable=partitionable)
        xn, r = pendulum_step(x[None], a, PendulumParams())
        x = xn[0]
        xs.append(x)
        rs.append(r[0])
        us.append(a[0])
    return np.array(xs), np.array(rs), np.array(us)


# ----------------------------------------------------------------------------------
# Vmapped env step with Episode/AutoReset wrappers and actor_step Transition
#   brax_wrapper.py:40-50; brax_utils/training.py:71-74,91-107,119-137; sac/acting.py:35-55
# ----------------------------------------------------------------------------------
def env_rollout(x0: np.ndarray, actions: np.ndarray, episode_length: int,
                p: PendulumParams = PendulumParams(), action_repeat: int = 1,
                steps0: Optional[np.ndarray] = None, done0: Optional[np.ndarray] = None,
                first_obs: Optional[np.ndarray] = None, dtype=F32):
    """x0[E,3], actions[T,E] -> dict of Transition fields, time-major [T,E,...].

    Per env and step: (1) AutoReset: steps = where(done, 0, steps); done = 0
    (training.py:120-124); (2) Episode: action_repeat x system.step with the same action,
    rewards summed (:92-97); (3) steps += action_repeat; done = where(steps >= L, 1, done);
    truncation = where(steps >= L, 1 - done_sys, 0) (:98-107); (4) obs = where(done,
    first_obs, obs) (:136); (5) Transition(observation=previous obs, action, reward,
    discount = 1 - done, next_observation = post-reset obs, truncation) (acting.py:46-55).
    """
    t_len, e = actions.shape
    obs = np.asarray(x0, dtype=dtype).copy()
    first = obs.copy() if first_obs is None else np.asarray(first_obs, dtype=dtype)
    steps = np.zeros(e, dtype=dtype) if steps0 is None else np.asarray(steps0, dtype=dtype).copy()
    done = np.zeros(e, dtype=dtype) if done0 is None else np.asarray(done0, dtype=dtype).copy()
    out = dict(observation=np.empty((t_len, e, 3), dtype), action=np.asarray(actions, dtype=dtype)[..., None],
               reward=np.empty((t_len, e), dtype), discount=np.empty((t_len, e), dtype),
               next_observation=np.empty((t_len, e, 3), dtype), truncation=np.empty((t_len, e), dtype))
    for t in range(t_len):
        steps = np.where(done != 0, dtype(0), steps)
        done = np.zeros(e, dtype=dtype)                               # system done is always 0.0
        prev = obs
        rew = np.zeros(e, dtype=dtype)
        x = obs
        for _ in range(action_repeat):
            x, r = pendulum_step(x, actions[t], p, dtype)
            rew = (rew + r).astype(dtype)
        steps = (steps + dtype(action_repeat)).astype(dtype)
        over = steps >= dtype(episode_length)
        trunc = np.where(over, dtype(1) - done, dtype(0)).astype(dtype)
        done = np.where(over, dtype(1), done).astype(dtype)
        obs = np.where(done[:, None] != 0, first, x).astype(dtype)
        out["observation"][t] = prev
        out["reward"][t] = rew
        out["discount"][t] = dtype(1) - done
        out["next_observation"][t] = obs
        out["truncation"][t] = trunc
    out["final_obs"] = obs
    out["final_steps"] = steps
    out["final_done"] = done
    return out


# ----------------------------------------------------------------------------------
# Policy in the env loop: actor_step / generate_unroll / SAC get_experience
#   sac/sac_networks.py:58-73 (make_inference_fn), sac/parametric_distribution.py:97-125
#   (NormalTanhDistribution), sac/acting.py:35-78, sac/sac.py:283-292
# ----------------------------------------------------------------------------------
@dataclass
class PolicyParams:
    """flax Dense kernels [in, out] and biases of the policy MLP (swish): obs -> hidden.. -> 2 * A."""
    weights: list
    biases: list
    min_std: float = 0.001


def make_policy_params(seed: int = 7, obs_dim: int = 3, action_dim: int = 1, hidden=(64, 64, 64),
                       scale: float = 1.0) -> PolicyParams:
    rng = np.random.default_rng(seed)
    dims = (obs_dim,) + tuple(hidden) + (2 * action_dim,)
    ws = [(rng.standard_normal((dims[i], dims[i + 1])) * scale / np.sqrt(dims[i])).astype(F32)
          for i in range(len(dims) - 1)]
    bs = [(0.1 * rng.standard_normal(dims[i + 1])).astype(F32) for i in range(len(dims) - 1)]
    return PolicyParams(ws, bs)


def policy_logits(params: PolicyParams, obs: np.ndarray) -> np.ndarray:
    """MLP(obs): Dense (x @ W + b) -> swish ... -> Dense, float32 (products summed in float64, rounded once)."""
    a = np.asarray(obs, dtype=F32)
    n = len(params.weights)
    for i in range(n):
        a = ((a.astype(np.float64) @ params.weights[i].astype(np.float64)).astype(F32) + params.biases[i]).astype(F32)
        if i < n - 1:
            a = swish(a.astype(F32)).astype(F32)
    return a


def softplus(x):
    """jax.nn.softplus = logaddexp(x, 0)."""
    x = np.asarray(x, dtype=F32)
    return (np.maximum(x, F32(0)) + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def policy_sample(params: PolicyParams, obs: np.ndarray, key, deterministic: bool = False,
                  partitionable: bool = False, obs_mean=None, obs_std=None, with_extras: bool = False):
    """make_policy(params)(observations, key_sample) (sac_networks.py:61-70): actions [E, A].  obs_mean / obs_std:
    the normaliser params applied by the policy network ((obs - mean) / std).  with_extras: PPO's policy
    (ppo/ppo_network.py:66-80) -> (actions, raw_actions [E, A], log_prob [E])."""
    x = np.asarray(obs, dtype=F32)
    if obs_mean is not None:
        x = ((x - np.asarray(obs_mean, F32)).astype(F32) / np.asarray(obs_std, F32)).astype(F32)
    logits = policy_logits(params, x)
    a_dim = logits.shape[-1] // 2
    loc, scale = logits[..., :a_dim], logits[..., a_dim:]              # jnp.split(parameters, 2, axis=-1)
    if deterministic:
        return np.tanh(loc).astype(F32)                                # mode(): postprocess(dist.mode())
    scale = (softplus(scale) + F32(params.min_std)).astype(F32)
    e = obs.shape[0]
    eps = jr.normal(key, e * a_dim, partitionable).reshape(e, a_dim)   # Normal._sample_n: normal(key, (1, E, A))
    raw = ((scale * eps).astype(F32) + loc).astype(F32)                # scale * rnd + loc
    act = np.tanh(raw).astype(F32)                                     # distrax.Tanh().forward
    if not with_extras:
        return act
    # parametric_distribution.py:76-83 with distrax 0.1.x: Normal.log_prob = -0.5 * ((x - loc) / scale)^2 -
    # (0.5 * log(2 pi) + log(scale)); Tanh.forward_log_det_jacobian(x) = 2 * (log(2) - x - softplus(-2 x))
    z = ((raw - loc).astype(F32) / scale).astype(F32)
    lp = (F32(-0.5) * (z * z).astype(F32) - (F32(0.5 * np.log(2 * np.pi)) + np.log(scale).astype(F32))).astype(F32)
    ldj = (F32(2.0) * ((F32(np.log(2.0)) - raw).astype(F32) - softplus((F32(-2.0) * raw).astype(F32)))).astype(F32)
    return act, raw, (lp - ldj).astype(F32).sum(axis=-1)


def actor_rollout(params: PolicyParams, x0: np.ndarray, key, num_steps: int, episode_length: int,
                  key_convention: str = "sac", deterministic: bool = False, p: PendulumParams = PendulumParams(),
                  action_repeat: int = 1, partitionable: bool = False, teacher_obs: Optional[np.ndarray] = None):
    """T steps of actor_step inside the reference's scans.  key_convention: "sac" (sac.py:288-292:
    k, k_t = split(k), policy key k_t), "unroll" (acting.py:68-73: current, next = split(current), policy key
    current, carry next) or "as_is" (one actor_step with the given key).  teacher_obs [T, E, 3] forces the
    observation seen at step t (per-step parity without chaotic drift).  Returns (Transition dict, carry key)."""
    e = x0.shape[0]
    obs = np.asarray(x0, dtype=F32).copy()
    first = obs.copy()
    steps = np.zeros(e, F32)
    done = np.zeros(e, F32)
    out = dict(observation=np.empty((num_steps, e, 3), F32), action=np.empty((num_steps, e, 1), F32),
               reward=np.empty((num_steps, e), F32), discount=np.empty((num_steps, e), F32),
               next_observation=np.empty((num_steps, e, 3), F32), truncation=np.empty((num_steps, e), F32))
    key = np.asarray(key, dtype=U32)
    for t in range(num_steps):
        if key_convention == "as_is":
            k_actor = key
        else:
            ks = jr.split(key, 2, partitionable)
            if key_convention == "sac":
                key, k_actor = ks[0], ks[1]
            else:
                k_actor, key = ks[0], ks[1]
        if teacher_obs is not None:
            obs = np.asarray(teacher_obs[t], dtype=F32)
        act = policy_sample(params, obs, k_actor, deterministic, partitionable)
        one = env_rollout(obs, act[:, 0][None], episode_length, p, action_repeat, steps0=steps, done0=done,
                          first_obs=first)
        out["observation"][t] = obs
        out["action"][t] = act
        for f in ("reward", "discount", "next_observation", "truncation"):
            out[f][t] = one[f][0]
        obs = one["next_observation"][0]
        steps, done = one["final_steps"], one["final_done"]
    return out, key


# ----------------------------------------------------------------------------------
# rollout_policy + its cotangent pass, lambda_return (BPTT)
#   mbpo/utils/optimizer_utils.py:62-131; bptt_optimizer.py:123-142 (Actor), :306-326 (act),
#   :327-352 (actor_loss), :361-376 (value_and_grad through the rollout)
# ----------------------------------------------------------------------------------
@dataclass
class BpttActorParams:
    """The BPTT actor: MLP params (PolicyParams layout), Actor's init_stddev / sig_min / sig_max
    (bptt_optimizer.py:127-129) and the state normaliser's mean / std (:70-72)."""
    mlp: PolicyParams
    init_stddev: float = 1.0
    sig_min: float = 1e-6
    sig_max: float = 1e2
    obs_mean: Optional[np.ndarray] = None
    obs_std: Optional[np.ndarray] = None


def inv_softplus(x: float) -> np.float32:
    """bptt_optimizer.py:107-108: where(x < 20, log(exp(x) - 1), x) on a weak-typed python float."""
    x32 = F32(x)
    return F32(np.log(np.exp(x32) - F32(1.0))) if x < 20.0 else x32


def bptt_actor(params: BpttActorParams, obs: np.ndarray):
    """Actor.__call__ on normalised observations: (mu, sig) each [E, A]."""
    x = np.asarray(obs, dtype=F32)
    if params.obs_mean is not None:
        x = ((x - np.asarray(params.obs_mean, F32)).astype(F32) / np.asarray(params.obs_std, F32)).astype(F32)
    out = policy_logits(params.mlp, x)
    a_dim = out.shape[-1] // 2
    mu, sig = out[..., :a_dim], out[..., a_dim:]
    sig = softplus((sig + inv_softplus(params.init_stddev)).astype(F32))
    sig = np.clip(sig, F32(params.sig_min), F32(params.sig_max)).astype(F32)
    return mu.astype(F32), sig


def bptt_act(params: BpttActorParams, obs: np.ndarray, key, evaluate: bool, partitionable: bool = False):
    """BPTT.act (:306-326) for a batch of observations sharing one key (the key is not vmapped at :366-368):
    returns (actions [E, A], new key)."""
    mu, sig = bptt_actor(params, obs)
    squash = lambda v: np.clip(np.tanh(v).astype(F32), F32(-0.999), F32(0.999)).astype(F32)
    if evaluate:
        return squash(mu), np.asarray(key, dtype=U32)
    ks = jr.split(key, 2, partitionable)
    sample_key, new_key = ks[0], ks[1]
    eps = jr.normal(sample_key, mu.shape[-1], partitionable)          # normal(sample_key, mu.shape), mu [A] per trajectory
    return squash((mu + (eps[None, :] * sig).astype(F32)).astype(F32)), new_key


def rollout_policy(params: BpttActorParams, x0: np.ndarray, key, horizon: int, evaluate: bool = False,
                   p: PendulumParams = PendulumParams(), partitionable: bool = False,
                   teacher_obs: Optional[np.ndarray] = None):
    """vmap(rollout_policy, in_axes=(init_state: 0, policy_state: None)) (optimizer_utils.py:62-116) with BPTT's
    train_policy: per step acs, new_state = policy(obs, state); system.step.  x0 [B, 3] -> dict of [B, H, ...]
    arrays (observation, action, reward, next_observation, discount) and the carried key."""
    b = x0.shape[0]
    obs = np.asarray(x0, dtype=F32).copy()
    out = dict(observation=np.empty((b, horizon, 3), F32), action=np.empty((b, horizon, 1), F32),
               reward=np.empty((b, horizon), F32), next_observation=np.empty((b, horizon, 3), F32),
               discount=np.ones((b, horizon), F32))
    key = np.asarray(key, dtype=U32)
    for t in range(horizon):
        if teacher_obs is not None:
            obs = np.asarray(teacher_obs[:, t], dtype=F32)
        act, key = bptt_act(params, obs, key, evaluate, partitionable)
        nxt, r = pendulum_step(obs, act[:, 0], p)
        out["observation"][:, t] = obs
        out["action"][:, t] = act
        out["reward"][:, t] = r
        out["next_observation"][:, t] = nxt
        obs = nxt
    return out, key


def _clip_grad(x, lo, hi):
    """d clip(x, lo, hi) / dx for jnp.clip = minimum(maximum(x, lo), hi); lax.max / lax.min give 0.5 at a tie."""
    a = np.where(x > lo, 1.0, np.where(x == lo, 0.5, 0.0))
    m = np.maximum(x, lo)
    b = np.where(m < hi, 1.0, np.where(m == hi, 0.5, 0.0))
    return a * b


def pendulum_step_vjp(x, u, g_next, g_reward, p: PendulumParams = PendulumParams(), dtype=F32):
    """Cotangents of one PendulumSystem.step: x [n,3], u [n], g_next [n,3], g_reward [n] -> (g_x [n,3], g_u [n]).
    Derived by hand from pendulum_dynamics.py:29-63 and pendulum_reward.py:27-42; checked against central
    differences of ``pendulum_step`` in float64 (tests/test_oracle_icem.py)."""
    x = np.asarray(x, dtype=dtype); u = np.asarray(u, dtype=dtype)
    g_next = np.asarray(g_next, dtype=dtype); g_reward = np.asarray(g_reward, dtype=dtype)
    c, s, w = x[:, 0], x[:, 1], x[:, 2]
    th = np.arctan2(s, c)
    c_g = dtype(3.0) * dtype(p.g) / (dtype(2.0) * dtype(p.l))
    c_u = dtype(3.0) / (dtype(p.m) * dtype(p.l) ** 2)
    uu = np.clip(u, dtype(-1), dtype(1)) * dtype(p.max_torque)
    thdd = c_g * np.sin(th) + c_u * uu
    v = w + thdd * dtype(p.dt)
    nw = np.clip(v, dtype(-p.max_speed), dtype(p.max_speed))
    nth = th + nw * dtype(p.dt)
    g_nth = np.cos(nth) * g_next[:, 1] - np.sin(nth) * g_next[:, 0]
    g_nw = g_next[:, 2] + dtype(p.dt) * g_nth
    g_v = g_nw * _clip_grad(v, dtype(-p.max_speed), dtype(p.max_speed)).astype(dtype)
    g_thdd = dtype(p.dt) * g_v
    two_pi = dtype(2 * np.pi)
    diff = np.mod(th - dtype(p.target_angle) + dtype(np.pi), two_pi) - dtype(np.pi)
    g_th = g_nth + g_thdd * c_g * np.cos(th) - g_reward * dtype(2.0) * dtype(p.angle_cost) * diff
    g_w = g_v - g_reward * dtype(0.2) * w
    g_u = (g_thdd * c_u * dtype(p.max_torque) * _clip_grad(u, dtype(-1), dtype(1)).astype(dtype)
           - g_reward * dtype(2.0) * dtype(p.control_cost) * u)
    r2 = c * c + s * s
    inv = np.where(r2 > 0, dtype(1) / np.where(r2 > 0, r2, dtype(1)), dtype(0))
    g_x = np.stack([-g_th * s * inv, g_th * c * inv, g_w], axis=-1)
    return g_x.astype(dtype), g_u.astype(dtype)


def rollout_policy_vjp(observation, action, g_reward=None, g_next_obs=None, g_obs=None, g_action=None,
                       p: PendulumParams = PendulumParams(), dtype=F32):
    """Cotangent pass of jax.grad through rollout_policy(..., stop_grads=True): observation [B,H,3], action
    [B,H,1] and the cotangents of the Transition fields -> (g_action_total [B,H,1], g_x0 [B,3]).
    g_action_total[t] is what reaches a_t = policy(stop_gradient(obs_t)); the parameter gradient is
    sum_t (da_t / dtheta)^T g_action_total[t]."""
    b, h = observation.shape[:2]
    z1, z3 = np.zeros((b, h), dtype), np.zeros((b, h, 3), dtype)
    g_reward = z1 if g_reward is None else np.asarray(g_reward, dtype)
    g_next_obs = z3 if g_next_obs is None else np.asarray(g_next_obs, dtype)
    g_obs = z3 if g_obs is None else np.asarray(g_obs, dtype)
    g_action = np.zeros((b, h, 1), dtype) if g_action is None else np.asarray(g_action, dtype)
    lam = np.zeros((b, 3), dtype)
    out = np.empty((b, h, 1), dtype)
    for t in range(h - 1, -1, -1):
        g_next = lam + g_next_obs[:, t] + (g_obs[:, t + 1] if t + 1 < h else 0)
        lam, g_u = pendulum_step_vjp(observation[:, t], action[:, t, 0], g_next, g_reward[:, t], p, dtype)
        out[:, t, 0] = g_u + g_action[:, t, 0]
    return out, (lam + (g_obs[:, 0] if h > 0 else 0)).astype(dtype)


def lambda_return(reward, next_values, discount: float, lambda_: float, dtype=F32):
    """optimizer_utils.py:119-131 along the last axis ([..., H]); python floats are weakly typed."""
    reward = np.asarray(reward, dtype); next_values = np.asarray(next_values, dtype)
    inputs = (reward + ((dtype(discount) * next_values).astype(dtype) * dtype(1 - lambda_)).astype(dtype)).astype(dtype)
    dl = dtype(discount * lambda_)
    agg = next_values[..., -1]
    out = np.empty_like(inputs)
    for t in range(inputs.shape[-1] - 1, -1, -1):
        agg = (inputs[..., t] + (dl * agg).astype(dtype)).astype(dtype)
        out[..., t] = agg
    return out


def lambda_return_vjp(g_returns, discount: float, lambda_: float, dtype=F32):
    """Transpose of lambda_return: (g_reward, g_next_values), each [..., H]."""
    g = np.asarray(g_returns, dtype)
    dl, dn = dtype(discount * lambda_), dtype(discount) * dtype(1 - lambda_)
    g_inp = np.empty_like(g)
    acc = np.zeros(g.shape[:-1], dtype)
    for t in range(g.shape[-1]):
        acc = (g[..., t] + dl * acc).astype(dtype)
        g_inp[..., t] = acc
    g_nv = (dn * g_inp).astype(dtype)
    g_nv[..., -1] += dl * g_inp[..., -1]
    return g_inp, g_nv


# ----------------------------------------------------------------------------------
# iCEM generality: Systems that consume the per-particle key and A > 1.
# The reference ships neither (its only System is the deterministic pendulum, whose
# PendulumDynamics.next_state already returns distrax.Normal(mean, std = 0),
# pendulum_dynamics.py:45-46); these two are the smallest Systems that exercise
# iCemTO's own code for them:
#   icem_optimizer.py:146-147  system_params.replace(key=rng)      one key per particle
#   icem_optimizer.py:155-156  split(key, num_particles) + vmap     P distinct rollouts
#   icem_optimizer.py:160      mean over the horizon, mean / max over particles
#   icem_optimizer.py:180      vmap(split(x, action_dim))           one noise key per action dim
#   optimizer_utils.py:28-46   the scan carries [obs, system_params]: the key threads through
# ----------------------------------------------------------------------------------
class NoisyPendulumOracle:
    """PendulumSystem whose transition is SAMPLED: next_state returns Normal(mean, noise_std) and step draws from it
    with the System's own key -- ``key, sub = split(system_params.key)``; ``x_next = mean + noise_std *
    normal(sub, (3,))`` (distrax Normal.sample: loc + scale * jax.random.normal) -- and carries ``key`` on in the
    returned SystemParams.  The reward is the pendulum reward on the current (noisy) state."""
    x_dim, u_dim, keyed = 3, 1, True

    def __init__(self, noise_std: float = 0.05, p: PendulumParams = PendulumParams()):
        self.noise_std, self.p = noise_std, p

    def step(self, x, u, keys, partitionable=False, dtype=F32):
        """x [R,3], u [R,1], keys uint32 [R,2] -> (x_next [R,3], reward [R], keys_next [R,2])."""
        mean, reward = pendulum_step(x, u[:, 0], self.p, dtype)
        ks = split_keys(keys, 2, partitionable)
        z = jr.bits_to_normal(random_bits_keys(ks[:, 1], 3, partitionable)).astype(dtype)
        x_next = (mean + (dtype(self.noise_std) * z).astype(dtype)).astype(dtype)
        return x_next, reward, ks[:, 0]


@dataclass(frozen=True)
class PointMassParams:
    dt: float = 0.1
    max_accel: float = 1.0
    max_speed: float = 2.0
    target_x: float = 1.0
    target_y: float = -0.5
    speed_cost: float = 0.1
    control_cost: float = 0.02

    def packed(self) -> np.ndarray:
        return np.array([self.dt, self.max_accel, self.max_speed, self.target_x, self.target_y, self.speed_cost,
                         self.control_cost], dtype=F32)


class PointMassOracle:
    """A planar double integrator with TWO action dimensions: state [px, py, vx, vy], action [ax, ay].
        a = clip(u, -1, 1) * max_accel;  v' = clip(v + a * dt, +-max_speed);  p' = p + v' * dt
        reward = -(|p - target|^2 + speed_cost * |v|^2) - control_cost * |u|^2     (on the current state, raw action)
    Deterministic; every float operation is a separate float32 rounding in the order written."""
    x_dim, u_dim, keyed = 4, 2, False

    def __init__(self, p: PointMassParams = PointMassParams()):
        self.p = p

    def step(self, x, u, keys=None, partitionable=False, dtype=F32):
        d, p = dtype, self.p
        x, u = np.asarray(x, d), np.asarray(u, d)
        pos, vel = x[:, 0:2], x[:, 2:4]
        tgt = np.array([p.target_x, p.target_y], d)
        dp = (pos - tgt).astype(d)
        dist2 = ((dp[:, 0] * dp[:, 0]).astype(d) + (dp[:, 1] * dp[:, 1]).astype(d)).astype(d)
        spd2 = ((vel[:, 0] * vel[:, 0]).astype(d) + (vel[:, 1] * vel[:, 1]).astype(d)).astype(d)
        u2 = ((u[:, 0] * u[:, 0]).astype(d) + (u[:, 1] * u[:, 1]).astype(d)).astype(d)
        reward = ((-(dist2 + (d(p.speed_cost) * spd2).astype(d)).astype(d)) - (d(p.control_cost) * u2).astype(d)).astype(d)
        a = (np.clip(u, d(-1), d(1)) * d(p.max_accel)).astype(d)
        nv = np.clip((vel + (a * d(p.dt)).astype(d)).astype(d), d(-p.max_speed), d(p.max_speed)).astype(d)
        npos = (pos + (nv * d(p.dt)).astype(d)).astype(d)
        return np.concatenate([npos, nv], axis=1).astype(d), reward, keys


class PendulumOracle:
    """The reference's deterministic pendulum behind the same interface (the key is dropped: pendulum_system.py:38)."""
    x_dim, u_dim, keyed = 3, 1, False

    def __init__(self, p: PendulumParams = PendulumParams()):
        self.p = p

    def step(self, x, u, keys=None, partitionable=False, dtype=F32):
        xn, r = pendulum_step(x, u[:, 0], self.p, dtype)
        return xn, r, keys


def system_rollout_actions(system, x0, actions, keys=None, partitionable=False, dtype=F32, full=False):
    """vmap(rollout_actions) (optimizer_utils.py:11-59) for any oracle System: x0 [R,X], actions [R,H,A], keys
    uint32 [R,2] (keyed Systems) -> horizon-mean reward [R] (+ obs [R,H,X], reward [R,H], next_obs [R,H,X])."""
    actions = np.asarray(actions, dtype)
    r, h, _ = actions.shape
    x = np.broadcast_to(np.asarray(x0, dtype), (r, system.x_dim)).copy()
    k = None if keys is None else np.asarray(keys, U32).reshape(r, 2).copy()
    acc = np.zeros(r, dtype)
    obs = np.empty((r, h, system.x_dim), dtype); nxt = np.empty_like(obs); rew = np.empty((r, h), dtype)
    for t in range(h):
        xn, rw, k = system.step(x, actions[:, t], k, partitionable, dtype)
        obs[:, t], nxt[:, t], rew[:, t] = x, xn, rw
        acc = (acc + rw).astype(dtype)
        x = xn
    ret = (acc / dtype(h)).astype(dtype)
    return (ret, obs, rew, nxt) if full else ret


def system_objective(system, x0, acts, particle_keys, params: ICemParams, use_optimism=False, partitionable=False,
                     dtype=F32):
    """vmap(objective) (icem_optimizer.py:144-160,195) for any oracle System: acts [M,H,A], particle_keys [M,2].
    Per candidate ``split(key, P)``, one rollout per particle key, mean over the horizon, then mean (left-to-right)
    or max over the particles."""
    m, P = acts.shape[0], params.num_particles
    if not system.keyed:
        ret = system_rollout_actions(system, x0, acts, None, partitionable, dtype)
        return ret if (use_optimism or P == 1) else particle_mean(np.repeat(ret[:, None], P, axis=1), dtype)
    pk = split_keys(particle_keys, P, partitionable)                              # :155  [M, P, 2]
    per = np.stack([system_rollout_actions(system, x0, acts, pk[:, p], partitionable, dtype) for p in range(P)], axis=1)
    return per.max(axis=1) if use_optimism else particle_mean(per, dtype)


def system_icem_optimize(system, x0, state: ICemState, params: ICemParams, horizon: int, use_optimism=False,
                         partitionable=False, dtype=F32, trace: Optional[list] = None) -> ICemState:
    """iCemTO.optimize (icem_optimizer.py:134-252) over any oracle System (action_dim = system.u_dim)."""
    A = system.u_dim
    mean = np.zeros((horizon, A), dtype)
    if params.warm_start:
        mean[:-1] = state.best_sequence[1:]
        mean[-1] = state.best_sequence[-1]
    std = np.full((horizon, A), params.init_std, dtype)
    best_seq, best_value = mean.copy(), dtype(-np.inf)
    ks = jr.split(state.key, 2, partitionable)
    carry_key, new_state_key = ks[0], ks[1]
    for _ in range(params.num_steps):
        carry_key, acts, pkeys = icem_sample_actions(carry_key, mean, std, params, horizon, A, partitionable, dtype)
        values = system_objective(system, x0, acts, pkeys, params, use_optimism, partitionable, dtype)
        mean, std, best_value, best_seq, idx = icem_refit(acts, values, mean, std, best_value, best_seq, params, dtype)
        if trace is not None:
            trace.append(dict(actions=acts, values=values, elite_idx=idx, mean=mean, std=std, best_value=best_value,
                              particle_keys=pkeys))
    return ICemState(key=new_state_key, best_sequence=np.asarray(best_seq, F32), best_reward=F32(best_value))
