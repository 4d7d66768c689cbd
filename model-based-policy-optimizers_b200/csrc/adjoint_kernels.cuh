// Reverse-mode pass through System.step rollouts, and the lambda-return scan.
//
// BPTT differentiates rollout_policy (mbpo/utils/optimizer_utils.py:62-116) with
// jax.value_and_grad (bptt_optimizer.py:361-376).  With stop_grads=True (:337, the only way the
// reference calls it) the policy sees stop_gradient(obs), so the cotangent reaches the policy
// parameters only through the actions:
//     x_{t+1} = f(x_t, a_t)        r_t = r(x_t, a_t)        a_t = pi_theta(sg(x_t))
//     lam_T   = 0
//     lam_t   = (df/dx)^T (g_next_obs[t] + lam_{t+1}) + g_reward[t] dr/dx          (state adjoint)
//     ga_t    = g_action[t] + (df/da)^T (g_next_obs[t] + lam_{t+1}) + g_reward[t] dr/da
// rollout_adjoint_pendulum_kernel runs that reverse scan for one trajectory per thread (the 3 x 3
// Jacobian-transpose products written out for pendulum_dynamics.py:29-63 and pendulum_reward.py:27-42),
// re-deriving the step's intermediates from the stored (observation, action) instead of saving them.
// The parameter gradient sum_t (da_t/dtheta)^T ga_t has no sequential dependency left and is one
// batched network backward over all T*E (obs, ga) rows -- the caller's autodiff.
//
// lambda_return_kernel: optimizer_utils.py:119-131 (Dreamer's lambda return): inputs_t = reward_t +
// discount * next_values_t * (1 - lambda); returns_t = inputs_t + discount * lambda * returns_{t+1},
// returns_T = next_values[-1] -- and its transpose (a forward scan), since BPTT differentiates it too.
//
// All arrays are addressed as base[t * stride_t + e * stride_e] (elements), so both the vmapped
// reference layout [B, H, ...] and the time-major [T, E, ...] layout of the rollout kernels work
// without copies; time-major makes every warp access a contiguous line.
#pragma once
#include "pendulum.cuh"

namespace mbpo {

struct AdjointArgs {
  MbpoPendulumParams sys;
  int E, T;
  long long st_t, st_e;            // strides of the scalar-per-step arrays (action, reward, g_*), in elements
  long long sx_t, sx_e;            // strides of the [.,.,3] arrays, in elements (innermost stride 1)
  const float* observation;        // x_t
  const float* action;             // a_t
  const float* g_reward;           // cotangent of reward[t]            (NULL = 0)
  const float* g_next_obs;         // cotangent of next_observation[t]  (NULL = 0)
  const float* g_obs;              // cotangent of observation[t]       (NULL = 0; observation[t] = next_observation[t-1], observation[0] = x0)
  const float* g_action_in;        // direct cotangent of action[t]     (NULL = 0)
  float* g_action_out;             // total cotangent reaching a_t
  float* g_x0_out;                 // [E,3] cotangent of the initial state (NULL = not wanted)
};

// jnp.clip = minimum(maximum(x, lo), hi): lax.max / lax.min split the cotangent evenly at a tie.
__device__ __forceinline__ float clip_grad(float x, float lo, float hi) {
  const float a = (x > lo) ? 1.0f : (x == lo ? 0.5f : 0.0f);          // d max(x, lo) / dx
  const float m = fmaxf(x, lo);
  const float b = (m < hi) ? 1.0f : (m == hi ? 0.5f : 0.0f);          // d min(m, hi) / dm
  return a * b;
}

__global__ void __launch_bounds__(128) rollout_adjoint_pendulum_kernel(const __grid_constant__ AdjointArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.E) return;
  const PendulumConsts pc(a.sys);
  const long long be = static_cast<long long>(e) * a.st_e, bx = static_cast<long long>(e) * a.sx_e;
  float lc = 0.0f, ls = 0.0f, lw = 0.0f;     // lam_{t+1}
  for (int t = a.T - 1; t >= 0; --t) {
    const long long i1 = be + t * a.st_t, i3 = bx + t * a.sx_t;
    const float c = a.observation[i3], s = a.observation[i3 + 1], w = a.observation[i3 + 2];
    const float u = a.action[i1];
    const float gr = a.g_reward ? a.g_reward[i1] : 0.0f;
    float gc = lc, gs = ls, gw = lw;
    if (a.g_next_obs) { gc += a.g_next_obs[i3]; gs += a.g_next_obs[i3 + 1]; gw += a.g_next_obs[i3 + 2]; }
    if (a.g_obs && t + 1 < a.T) {            // observation[t+1] is the same value as next_observation[t]
      const long long j3 = i3 + a.sx_t;
      gc += a.g_obs[j3]; gs += a.g_obs[j3 + 1]; gw += a.g_obs[j3 + 2];
    }
    // forward intermediates of this step (pendulum_dynamics.py:35,59-62,40,43)
    const float th = atan2_bounded(s, c);
    float sin_th, cos_th;
    sincos_bounded(th, sin_th, cos_th);
    const float uu = __fmul_rn(fminf(fmaxf(u, -1.0f), 1.0f), pc.max_torque);
    const float thdd = fmaf(pc.c_g, sin_th, __fmul_rn(pc.c_u, uu));
    const float v = fmaf(thdd, pc.dt, w);
    const float nw = fminf(fmaxf(v, -pc.max_speed), pc.max_speed);
    const float nth = fmaf(nw, pc.dt, th);
    float sn, cn;
    sincos_bounded(nth, sn, cn);
    // x_next = [cos(nth), sin(nth), nw]
    const float g_nth = cn * gs - sn * gc;
    const float g_nw = gw + pc.dt * g_nth;
    const float g_v = g_nw * clip_grad(v, -pc.max_speed, pc.max_speed);
    const float g_thdd = pc.dt * g_v;
    // reward on (x_t, a_t) (pendulum_reward.py:32-39); d/dth of the wrapped difference is 1
    const float diff = wrap_diff(th - pc.target_angle);
    float g_th = g_nth + g_thdd * pc.c_g * cos_th - gr * 2.0f * pc.angle_cost * diff;
    const float g_w = g_v - gr * 0.2f * w;
    float g_u = g_thdd * pc.c_u * pc.max_torque * clip_grad(u, -1.0f, 1.0f) - gr * 2.0f * pc.control_cost * u;
    if (a.g_action_in) g_u += a.g_action_in[i1];
    a.g_action_out[i1] = g_u;
    // th = atan2(s, c): dth/dc = -s / (c^2 + s^2), dth/ds = c / (c^2 + s^2)
    const float r2 = c * c + s * s;
    const float inv = r2 > 0.0f ? 1.0f / r2 : 0.0f;
    lc = -g_th * s * inv;
    ls = g_th * c * inv;
    lw = g_w;
  }
  if (a.g_x0_out) {
    float gc = lc, gs = ls, gw = lw;
    if (a.g_obs && a.T > 0) { gc += a.g_obs[bx]; gs += a.g_obs[bx + 1]; gw += a.g_obs[bx + 2]; }
    a.g_x0_out[3 * e] = gc; a.g_x0_out[3 * e + 1] = gs; a.g_x0_out[3 * e + 2] = gw;
  }
}

struct LambdaArgs {
  int E, T;
  long long st_t, st_e;
  float discount, lambda_, one_minus_lambda, discount_lambda;
  const float* reward;        // forward: reward;        transpose: g_returns
  const float* next_values;   // forward: next_values;   transpose: unused
  float* out;                 // forward: returns;       transpose: g_reward
  float* out2;                // forward: unused;        transpose: g_next_values
};

// returns[t] = reward[t] + discount * next_values[t] * (1 - lambda) + discount * lambda * returns[t+1]
__global__ void __launch_bounds__(128) lambda_return_kernel(const __grid_constant__ LambdaArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.E || a.T == 0) return;
  const long long be = static_cast<long long>(e) * a.st_e;
  const float one_minus = a.one_minus_lambda, dl = a.discount_lambda;
  float agg = a.next_values[be + (a.T - 1) * a.st_t];
  for (int t = a.T - 1; t >= 0; --t) {
    const long long i = be + t * a.st_t;
    // inputs = reward + discount * next_values * (1 - lambda_)   (left to right, unfused)
    const float inp = __fadd_rn(a.reward[i], __fmul_rn(__fmul_rn(a.discount, a.next_values[i]), one_minus));
    agg = __fadd_rn(inp, __fmul_rn(dl, agg));        // inp + discount * lambda_ * agg
    a.out[i] = agg;
  }
}

// Transpose of the scan above: g_inputs[t] = g_returns[t] + discount * lambda * g_inputs[t-1];
// g_reward = g_inputs; g_next_values = discount * (1 - lambda) * g_inputs, plus the bootstrap
// start next_values[-1], which receives discount * lambda * g_inputs[T-1].
__global__ void __launch_bounds__(128) lambda_return_transpose_kernel(const __grid_constant__ LambdaArgs a) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.E || a.T == 0) return;
  const long long be = static_cast<long long>(e) * a.st_e;
  const float dl = a.discount_lambda, dn = a.discount * a.one_minus_lambda;
  float g = 0.0f;
  for (int t = 0; t < a.T; ++t) {
    const long long i = be + t * a.st_t;
    g = fmaf(dl, g, a.reward[i]);
    a.out[i] = g;
    a.out2[i] = dn * g + (t == a.T - 1 ? dl * g : 0.0f);
  }
}

}  // namespace mbpo
