// Vmapped env rollouts for SAC/PPO data collection (BASELINE config 3).
//
// One thread per environment runs T wrapped env steps and streams brax Transition fields,
// time-major, to HBM.  Per env and step (reference):
//   AutoResetWrapper.step   brax_utils/training.py:119-137  steps = where(done,0,steps); done = 0;
//                                                           obs = where(done_after, first_obs, obs)
//   VmapWrapper.step        brax_utils/training.py:71-74    (this kernel's thread index)
//   EpisodeWrapper.step     brax_utils/training.py:91-107   action_repeat x system.step, reward sum,
//                                                           steps += repeat, done/truncation at episode_length
//   BraxWrapper.step        systems/brax_wrapper.py:40-50   system.step(obs, action, params)
//   actor_step              sac/acting.py:35-55             Transition(obs_prev, action, reward,
//                                                           1 - done, obs_after_reset, truncation)
// HBM traffic per transition (pendulum): 4 B action read + 36 B written (obs 12, next_obs 12,
// reward 4, discount 4, truncation 4); the Transition's `action` field aliases the input.
#pragma once
#include "pendulum.cuh"

namespace mbpo {

struct EnvArgs {
  MbpoPendulumParams sys;
  int E, T, episode_length, action_repeat;
  float* obs;            // [E,3] in/out
  float* steps;          // [E]   in/out
  float* done;           // [E]   in/out
  const float* first_obs;  // [E,3]
  const float* actions;  // [T,E]
  float* observation_out;       // [T,E,3]
  float* reward_out;            // [T,E]
  float* discount_out;          // [T,E]
  float* next_observation_out;  // [T,E,3]
  float* truncation_out;        // [T,E]
};

constexpr int ENV_CHUNK = 8;   // actions prefetched per thread (registers)

// Transposes a warp's 32 x 3 floats through shared memory so the [E,3] rows leave as three
// fully coalesced 128-byte stores.
__device__ __forceinline__ void warp_store3(float* tile, float* dst_warp, int lane, int valid_rows, float a, float b,
                                            float c) {
  tile[lane * 3] = a;
  tile[lane * 3 + 1] = b;
  tile[lane * 3 + 2] = c;
  __syncwarp();
  const int n = valid_rows * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int i = lane + 32 * k;
    if (i < n) dst_warp[i] = tile[i];
  }
  __syncwarp();
}

template <int MATH>
__global__ void __launch_bounds__(128) env_rollout_pendulum_kernel(const __grid_constant__ EnvArgs a) {
  __shared__ float tiles[4][96];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int warp_e0 = e - lane;
  if (warp_e0 >= a.E) return;
  const bool live = e < a.E;
  const int valid_rows = (a.E - warp_e0) < 32 ? (a.E - warp_e0) : 32;
  const int ee = live ? e : a.E - 1;  // dead lanes shadow the last env, their stores are masked
  const PendulumConsts pc(a.sys);
  float* tile = tiles[warp];

  float c = a.obs[3 * ee], s = a.obs[3 * ee + 1], w = a.obs[3 * ee + 2];
  const float f_c = a.first_obs[3 * ee], f_s = a.first_obs[3 * ee + 1], f_w = a.first_obs[3 * ee + 2];
  float steps = a.steps[ee], done = a.done[ee];
  const float ep_len = static_cast<float>(a.episode_length);
  const float rep = static_cast<float>(a.action_repeat);

  float u_buf[ENV_CHUNK];
#pragma unroll
  for (int k = 0; k < ENV_CHUNK; ++k) u_buf[k] = (k < a.T) ? __ldg(a.actions + static_cast<size_t>(k) * a.E + ee) : 0.0f;

  for (int t0 = 0; t0 < a.T; t0 += ENV_CHUNK) {
    float u_cur[ENV_CHUNK];
#pragma unroll
    for (int k = 0; k < ENV_CHUNK; ++k) u_cur[k] = u_buf[k];
    // prefetch the next chunk while this one computes
#pragma unroll
    for (int k = 0; k < ENV_CHUNK; ++k) {
      const int t = t0 + ENV_CHUNK + k;
      u_buf[k] = (t < a.T) ? __ldg(a.actions + static_cast<size_t>(t) * a.E + ee) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < ENV_CHUNK; ++k) {
      const int t = t0 + k;
      if (t >= a.T) break;
      const size_t row = static_cast<size_t>(t) * a.E;
      // AutoReset pre-step (training.py:120-124)
      steps = (done != 0.0f) ? 0.0f : steps;
      done = 0.0f;
      if (a.observation_out) warp_store3(tile, a.observation_out + (row + warp_e0) * 3, lane, valid_rows, c, s, w);
      // Episode: action_repeat x system.step with the same action (training.py:92-97)
      float rew = 0.0f;
      for (int r = 0; r < a.action_repeat; ++r) {
        float rr;
        if (MATH == MBPO_MATH_REFERENCE) {
          pendulum_step_ref(pc, c, s, w, u_cur[k], rr);
        } else {
          float th = atan2_bounded(s, c);
          pendulum_step_theta(pc, th, w, u_cur[k], rr);
          sincos_bounded(th, s, c);
        }
        rew = __fadd_rn(rew, rr);
      }
      steps = __fadd_rn(steps, rep);
      const bool over = steps >= ep_len;                   // training.py:98-107
      const float trunc = over ? (1.0f - done) : 0.0f;     // system done is always 0.0
      done = over ? 1.0f : done;
      if (done != 0.0f) { c = f_c; s = f_s; w = f_w; }     // training.py:136
      if (a.next_observation_out)
        warp_store3(tile, a.next_observation_out + (row + warp_e0) * 3, lane, valid_rows, c, s, w);
      if (live) {
        if (a.reward_out) a.reward_out[row + e] = rew;
        if (a.discount_out) a.discount_out[row + e] = 1.0f - done;
        if (a.truncation_out) a.truncation_out[row + e] = trunc;
      }
    }
  }
  if (live) {
    a.obs[3 * e] = c; a.obs[3 * e + 1] = s; a.obs[3 * e + 2] = w;
    a.steps[e] = steps;
    a.done[e] = done;
  }
}

}  // namespace mbpo
