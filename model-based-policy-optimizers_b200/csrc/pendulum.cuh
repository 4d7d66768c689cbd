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
__device__ __forceinline__ float wrap_diff(float d) {
  const float x = d + MBPO_PI_F;
  float r;
  if (fabsf(x) < 2.0f * MBPO_TWO_PI_F) {
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

__device__ __forceinline__ float reward_from(const PendulumConsts& p, float th, float thdot, float u) {
  const float diff = wrap_diff(th - p.target_angle);
  return -(p.angle_cost * (diff * diff) + 0.1f * (thdot * thdot)) - p.control_cost * (u * u);
}

// One reference-literal step on the [cos, sin, thdot] state.
__device__ __forceinline__ void pendulum_step_ref(const PendulumConsts& p, float& c, float& s, float& w,
                                                  float u, float& reward) {
  const float th = atan2_bounded(s, c);                             // dynamics :35 / reward :32
  reward = reward_from(p, th, w, u);                                // reward uses x, raw u
  const float uu = fminf(fmaxf(u, -1.0f), 1.0f) * p.max_torque;     // :59
  const float thdd = p.c_g * sin_bounded(th) + p.c_u * uu;          // :60
  const float nw = fminf(fmaxf(w + thdd * p.dt, -p.max_speed), p.max_speed);   // :61-62 (=:41-42)
  const float nth = th + nw * p.dt;                                 // :40
  sincos_bounded(nth, s, c);                                        // :43
  w = nw;
}

// Theta-carry variant: the state is (theta, thdot) with theta kept in (-pi, pi], which is
// what atan2(sin(newth), cos(newth)) returns up to rounding.
__device__ __forceinline__ void pendulum_step_theta(const PendulumConsts& p, float& th, float& w, float u,
                                                    float& reward) {
  reward = reward_from(p, th, w, u);
  const float uu = fminf(fmaxf(u, -1.0f), 1.0f) * p.max_torque;
  const float thdd = p.c_g * sin_bounded(th) + p.c_u * uu;
  const float nw = fminf(fmaxf(w + thdd * p.dt, -p.max_speed), p.max_speed);
  float nth = th + nw * p.dt;
  if (nth > MBPO_PI_F) nth -= MBPO_TWO_PI_F;
  if (nth < -MBPO_PI_F) nth += MBPO_TWO_PI_F;
  th = nth;
  w = nw;
}

}  // namespace mbpo
