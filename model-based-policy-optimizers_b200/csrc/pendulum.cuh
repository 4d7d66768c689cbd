// Analytic pendulum System.step, inlined for the rollout kernels.
//
// Follows mbpo/systems/pendulum_system.py:18-39,
//         mbpo/systems/dynamics/pendulum_dynamics.py:29-63 (next_state / ode) and
//         mbpo/systems/rewards/pendulum_reward.py:27-42 (reward on the CURRENT state).
#pragma once
#include "../../include/mbpo_b200.h"
#include "mathx.cuh"

namespace mbpo {

// Loop-invariant constants derived from MbpoPendulumParams exactly as the reference's
// float32 expressions evaluate them: 3*g/(2*l) and 3.0/(m*l**2) (pendulum_dynamics.py:60).
struct PendulumConsts {
  float max_speed, max_torque, dt, c_g, c_u;
  float control_cost, angle_cost, target_angle;
  __host__ __device__ explicit PendulumConsts(const MbpoPendulumParams& p) {
    max_speed = p.max_speed;
    max_torque = p.max_torque;
    dt = p.dt;
#ifdef __CUDA_ARCH__
    c_g = __fdiv_rn(__fmul_rn(3.0f, p.g), __fmul_rn(2.0f, p.l));
    c_u = __fdiv_rn(3.0f, __fmul_rn(p.m, __fmul_rn(p.l, p.l)));
#else
    c_g = (3.0f * p.g) / (2.0f * p.l);
    c_u = 3.0f / (p.m * (p.l * p.l));
#endif
    control_cost = p.control_cost;
    angle_cost = p.angle_cost;
    target_angle = p.target_angle;
  }
};

#define MBPO_PI_F 3.14159274f      /* float32(jnp.pi)   */
#define MBPO_TWO_PI_F 6.28318548f  /* float32(2*jnp.pi) */

// ((d + pi) % (2*pi)) - pi with jnp's floored remainder (pendulum_reward.py:35).
// SMALL: the caller guarantees |d + pi| < 4*pi (|theta| <= pi + max_speed*dt and a target angle
// the host checked), which drops the fmod slow path from the instruction stream.
template <bool SMALL = false>
__device__ __forceinline__ float wrap_diff(float d) {
  const float x = d + MBPO_PI_F;
  float r;
  if (SMALL || fabsf(x) < 2.0f * MBPO_TWO_PI_F) {
    // fmod is exact; for |x| < 4*pi it is x or x -/+ 2*pi (Sterbenz-exact subtraction)
    r = x;
    if (r >= MBPO_TWO_PI_F) r -= MBPO_TWO_PI_F;
    if (r <= -MBPO_TWO_PI_F) r += MBPO_TWO_PI_F;
  } else {
    r = fmodf(x, MBPO_TWO_PI_F);
  }
  if (r < 0.0f) r += MBPO_TWO_PI_F;
  return r - MBPO_PI_F;
}

template <bool SMALL = false>
__device__ __forceinline__ float reward_from(const PendulumConsts& p, float th, float thdot, float u) {
  // Every contraction is spelled out: a*b + c*d leaves the compiler a choice of which product to
  // fuse, and it chooses differently in different loops -- the kernels must agree bit for bit.
  const float diff = wrap_diff<SMALL>(th - p.target_angle);
  const float t = fmaf(0.1f, __fmul_rn(thdot, thdot), __fmul_rn(p.angle_cost, __fmul_rn(diff, diff)));
  return fmaf(-p.control_cost, __fmul_rn(u, u), -t);
}

// thdot + (3g/(2l) sin(th) + 3/(ml^2) u) dt, clipped (pendulum_dynamics.py:59-62)
__device__ __forceinline__ float pendulum_next_thdot(const PendulumConsts& p, float sin_th, float w, float u) {
  const float uu = __fmul_rn(fminf(fmaxf(u, -1.0f), 1.0f), p.max_torque);   // :59
  const float thdd = fmaf(p.c_g, sin_th, __fmul_rn(p.c_u, uu));               // :60
  return fminf(fmaxf(fmaf(thdd, p.dt, w), -p.max_speed), p.max_speed);        // :61-62 (=:41-42)
}

// One reference-literal step on the [cos, sin, thdot] state.
template <bool SMALL = false>
__device__ __forceinline__ void pendulum_step_ref(const PendulumConsts& p, float& c, float& s, float& w,
                                                  float u, float& reward) {
  const float th = atan2_bounded(s, c);                             // dynamics :35 / reward :32
  reward = reward_from<SMALL>(p, th, w, u);                         // reward uses x, raw u
  const float nw = pendulum_next_thdot(p, sin_bounded(th), w, u);
  const float nth = fmaf(nw, p.dt, th);                             // :40
  sincos_bounded(nth, s, c);                                        // :43
  w = nw;
}

// The same reference-literal step for a loop that only needs the rewards: theta = atan2(sin, cos) of the CURRENT
// state comes in, theta of the NEXT state goes out -- atan2 of the freshly computed [cos(newth), sin(newth)] as
// pendulum_dynamics.py:35 evaluates it at the next step, so every value has the bits pendulum_step_ref produces;
// the pair is a unit vector, which lets atan2 skip its zero / denormal guard.  The first theta of a rollout is
// atan2_bounded(s0, c0) (guarded: the caller's initial state is arbitrary).
template <bool SMALL = false>
__device__ __forceinline__ void pendulum_step_ref_th(const PendulumConsts& p, float& th, float& w, float u,
                                                     float& reward) {
  reward = reward_from<SMALL>(p, th, w, u);
  const float nw = pendulum_next_thdot(p, sin_bounded(th), w, u);
  const float nth = fmaf(nw, p.dt, th);
  float s, c;
  sincos_bounded(nth, s, c);
  th = atan2_bounded<true>(s, c);
  w = nw;
}

// The same step for a loop that also needs [cos, sin] of every state (the env kernels stream them out): theta of the
// current state in, the next state's [cos, sin] and its theta out.  Same bits as pendulum_step_ref.
template <bool SMALL = false>
__device__ __forceinline__ void pendulum_step_ref_thcs(const PendulumConsts& p, float& th, float& c, float& s,
                                                       float& w, float u, float& reward) {
  reward = reward_from<SMALL>(p, th, w, u);
  const float nw = pendulum_next_thdot(p, sin_bounded(th), w, u);
  const float nth = fmaf(nw, p.dt, th);
  sincos_bounded(nth, s, c);
  th = atan2_bounded<true>(s, c);
  w = nw;
}

// Theta-carry variant: the state is (theta, thdot) with theta kept in (-pi, pi], which is
// what atan2(sin(newth), cos(newth)) returns up to rounding.
template <bool SMALL = false>
__device__ __forceinline__ void pendulum_step_theta(const PendulumConsts& p, float& th, float& w, float u,
                                                    float& reward) {
  reward = reward_from<SMALL>(p, th, w, u);
  const float nw = pendulum_next_thdot(p, sin_bounded(th), w, u);
  float nth = fmaf(nw, p.dt, th);
  if (nth > MBPO_PI_F) nth -= MBPO_TWO_PI_F;
  if (nth < -MBPO_PI_F) nth += MBPO_TWO_PI_F;
  th = nth;
  w = nw;
}

}  // namespace mbpo
