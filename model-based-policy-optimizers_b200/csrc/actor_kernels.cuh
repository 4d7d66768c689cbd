// Policy-in-the-loop data collection: actor_step / generate_unroll / SAC get_experience.
//
// Reference (per env and step):
//   policy            sac/sac_networks.py:58-73     logits = MLP(obs)  (swish, Dense = x @ W + b,
//                                                   utils/network_utils.py:5-17 style stack)
//   NormalTanh        sac/parametric_distribution.py:97-125   loc, scale = split(logits, 2);
//                                                   scale = softplus(scale) + min_std;
//                                                   raw = scale * normal(key, [E, A]) + loc; action = tanh(raw)
//                                                   (deterministic: action = tanh(loc))
//   actor_step        sac/acting.py:35-55           nstate = env.step(state, action); Transition(obs, action,
//                                                   reward, 1 - done, next_obs, truncation)
//   key per step      sac/sac.py:288-292            k, k_t = split(k); actor key = k_t          (convention 0)
//                     sac/acting.py:68-73           current, next = split(current); actor key = current,
//                                                   carry = next                                  (convention 1)
//   env.step          the wrapper stack of env_kernels.cuh (AutoReset / Episode / BraxWrapper / System.step)
//
// One thread per environment, T steps per launch.  The network runs in float32 on the CUDA cores (the
// reference network is float32; operand rounding to bf16 would change the collected actions): the
// thread's current activations live in 64 registers, the weights are read from shared memory with
// warp-broadcast 128-bit loads (8 output units per pass: 2 loads feed 8 FFMAs), the next layer's
// activations pass through a column of shared memory private to the thread (no block barrier).
// The random draw is JAX's: normal(key, (E * A,))[e] costs one threefry block per env and step.
// CTAs are sized so that the envs spread evenly over the SMs in a single wave (one CTA per SM).
#pragma once
#include "env_kernels.cuh"
#include "mathx.cuh"
#include "threefry.cuh"

namespace mbpo {

constexpr int ACT_W = 64;             // hidden width
constexpr int ACT_MAX_HIDDEN = 4;     // hidden layers
constexpr int ACT_MAX_THREADS = 448;  // 14 warps: 64 + ~80 registers per thread fit 65,536 / 448 = 146

struct ActorArgs {
  MbpoPendulumParams sys;
  int E, T, episode_length, action_repeat;
  int num_hidden, deterministic, key_convention;  // 0: SAC get_experience, 1: generate_unroll, 2: use key as is
  float min_std;
  const float* w[ACT_MAX_HIDDEN + 1];  // [3,64], [64,64] x (num_hidden-1), [64,2]   (flax Dense kernels, [in, out])
  const float* b[ACT_MAX_HIDDEN + 1];
  const uint32_t* key_in;       // [2] device
  float* obs;            // [E,3] in/out
  float* steps;          // [E]   in/out
  float* done;           // [E]   in/out
  const float* first_obs;       // [E,3]
  float* action_out;            // [T,E]
  float* reward_out;            // [T,E]
  float* discount_out;          // [T,E]
  float* next_observation_out;  // [T,E,3]
  float* truncation_out;        // [T,E]
  uint32_t* key_out;            // [2] carry key after T steps
};

// jax.nn.swish(x) = x * sigmoid(x), sigmoid = 1 / (1 + exp(-x)).  ex2.approx (2^-22 relative) and rcp.approx
// (1 ulp) keep the activation within ~5e-7 relative of the float32 reference at 5 instructions; the accurate
// expf + IEEE division cost 20, i.e. a quarter of the whole kernel for the 192 hidden units.
__device__ __forceinline__ float swish_exact(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + __expf(-x)));
  return x * r;
}
// jax.nn.softplus(x) = logaddexp(x, 0) = max(x, 0) + log1p(exp(-|x|))
__device__ __forceinline__ float softplus_exact(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

// Packed float32 pair arithmetic (sm_100: FFMA2 takes a scalar and a register pair).  A 3-register FFMA issues
// every other cycle per scheduler; the packed form retires two FMAs per issue slot, IEEE-rounded per element.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// acc += s * w  (both elements)
__device__ __forceinline__ void fma2_scalar(f32x2_t& acc, float s, f32x2_t w) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(pack2(s, s)), "l"(w));
}

struct ActorSmem {
  // float offsets into dynamic shared memory
  int w[ACT_MAX_HIDDEN + 1], b[ACT_MAX_HIDDEN + 1], h;
};

template <int PRNG, int MATH>
__global__ void __launch_bounds__(ACT_MAX_THREADS, 1) actor_rollout_pendulum_kernel(const __grid_constant__ ActorArgs a,
                                                                                    const ActorSmem lay) {
  extern __shared__ __align__(16) float act_sm[];
  __shared__ float tiles[ACT_MAX_THREADS / 32][96];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
  // ---- policy parameters -> shared memory (once per launch) ----------------------------------------
  {
    const int L = a.num_hidden;
    for (int i = tid; i < 3 * ACT_W; i += nthr) act_sm[lay.w[0] + i] = a.w[0][i];
    for (int l = 1; l < L; ++l)
      for (int i = tid; i < ACT_W * ACT_W; i += nthr) act_sm[lay.w[l] + i] = a.w[l][i];
    for (int i = tid; i < ACT_W * 2; i += nthr) act_sm[lay.w[L] + i] = a.w[L][i];
    for (int l = 0; l < L; ++l)
      for (int i = tid; i < ACT_W; i += nthr) act_sm[lay.b[l] + i] = a.b[l][i];
    if (tid < 2) act_sm[lay.b[L] + tid] = a.b[L][tid];
  }
  __syncthreads();
  const int e = blockIdx.x * nthr + tid;
  const int warp_e0 = e - lane;
  if (warp_e0 >= a.E) return;
  const bool live = e < a.E;
  const int n_valid = ((a.E - warp_e0) < 32 ? (a.E - warp_e0) : 32) * 3;
  const int ee = live ? e : a.E - 1;
  const PendulumConsts pc(a.sys);
  float* tile = tiles[warp];
  float* h_col = act_sm + lay.h + tid;     // this thread's activation column: h_col[k * nthr]

  float c = a.obs[3 * ee], s = a.obs[3 * ee + 1], w = a.obs[3 * ee + 2];
  const float f_c = a.first_obs[3 * ee], f_s = a.first_obs[3 * ee + 1], f_w = a.first_obs[3 * ee + 2];
  float steps = a.steps[ee], done = a.done[ee];
  const float ep_len = static_cast<float>(a.episode_length);
  const float rep = static_cast<float>(a.action_repeat);
  const size_t E = static_cast<size_t>(a.E);
  float th = (MATH == MBPO_MATH_REFERENCE) ? 0.0f : atan2_bounded(s, c);
  const float f_th = (MATH == MBPO_MATH_REFERENCE) ? 0.0f : atan2_bounded(f_s, f_c);
  Key2 key{a.key_in[0], a.key_in[1]};

  float* p_act = a.action_out + ee;
  float* p_rew = a.reward_out + ee;
  float* p_dis = a.discount_out + ee;
  float* p_tru = a.truncation_out + ee;
  float* p_nxt = a.next_observation_out + static_cast<size_t>(warp_e0) * 3 + lane;

#pragma unroll 1
  for (int t = 0; t < a.T; ++t) {
    // ---- key plumbing ---------------------------------------------------------------------------------
    Key2 k_actor = key;
    if (a.key_convention != 2) {
      Key2 first, second;
      split2<PRNG>(key, first, second);
      if (a.key_convention == 0) { key = first; k_actor = second; }   // sac.py:290
      else { k_actor = first; key = second; }                          // acting.py:70
    }
    // ---- policy network, float32 ----------------------------------------------------------------------
    float h[ACT_W];
    {
      const float4* w0 = reinterpret_cast<const float4*>(act_sm + lay.w[0]);   // [3][64]
      const float4* b0 = reinterpret_cast<const float4*>(act_sm + lay.b[0]);
#pragma unroll
      for (int j4 = 0; j4 < ACT_W / 4; ++j4) {
        const float4 r0 = w0[j4], r1 = w0[ACT_W / 4 + j4], r2 = w0[2 * (ACT_W / 4) + j4], bb = b0[j4];
        h[4 * j4 + 0] = swish_exact(fmaf(w, r2.x, fmaf(s, r1.x, c * r0.x)) + bb.x);
        h[4 * j4 + 1] = swish_exact(fmaf(w, r2.y, fmaf(s, r1.y, c * r0.y)) + bb.y);
        h[4 * j4 + 2] = swish_exact(fmaf(w, r2.z, fmaf(s, r1.z, c * r0.z)) + bb.z);
        h[4 * j4 + 3] = swish_exact(fmaf(w, r2.w, fmaf(s, r1.w, c * r0.w)) + bb.w);
      }
    }
#pragma unroll 1
    for (int l = 1; l < a.num_hidden; ++l) {
      const float* wl = act_sm + lay.w[l];
      const float* bl = act_sm + lay.b[l];
#pragma unroll 1
      for (int jc = 0; jc < ACT_W / 8; ++jc) {
        f32x2_t acc2[4] = {0ull, 0ull, 0ull, 0ull};      // 8 output units as 4 packed pairs
        const float4* wrow = reinterpret_cast<const float4*>(wl + jc * 8);
#pragma unroll
        for (int k = 0; k < ACT_W; ++k) {
          const float4 wa = wrow[k * (ACT_W / 4)], wb = wrow[k * (ACT_W / 4) + 1];
          fma2_scalar(acc2[0], h[k], pack2(wa.x, wa.y));
          fma2_scalar(acc2[1], h[k], pack2(wa.z, wa.w));
          fma2_scalar(acc2[2], h[k], pack2(wb.x, wb.y));
          fma2_scalar(acc2[3], h[k], pack2(wb.z, wb.w));
        }
        float acc[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) unpack2(acc2[i], acc[2 * i], acc[2 * i + 1]);
        const float4 ba = *reinterpret_cast<const float4*>(bl + jc * 8);
        const float4 bb = *reinterpret_cast<const float4*>(bl + jc * 8 + 4);
        const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) h_col[(jc * 8 + i) * nthr] = swish_exact(acc[i] + bias[i]);
      }
#pragma unroll
      for (int k = 0; k < ACT_W; ++k) h[k] = h_col[k * nthr];
    }
    float loc, raw_scale;
    {
      const float2* wo = reinterpret_cast<const float2*>(act_sm + lay.w[a.num_hidden]);   // [64][2]
      f32x2_t o2 = 0ull;
#pragma unroll
      for (int k = 0; k < ACT_W; ++k) {
        const float2 v = wo[k];
        fma2_scalar(o2, h[k], pack2(v.x, v.y));
      }
      unpack2(o2, loc, raw_scale);
      loc += act_sm[lay.b[a.num_hidden]];
      raw_scale += act_sm[lay.b[a.num_hidden] + 1];
    }
    float u;
    if (a.deterministic) {
      u = tanhf(loc);                                                   // mode(): tanh(loc)
    } else {
      const float eps = bits_to_normal(random_bits_at<PRNG>(k_actor, static_cast<uint32_t>(a.E), static_cast<uint32_t>(ee)));
      const float scale = softplus_exact(raw_scale) + a.min_std;
      u = tanhf(__fadd_rn(__fmul_rn(scale, eps), loc));                 // distrax Normal.sample: scale * rnd + loc
    }
    // ---- wrapped env step (env_kernels.cuh) --------------------------------------------------------------
    steps = (done != 0.0f) ? 0.0f : steps;
    done = 0.0f;
    float rew = 0.0f;
    for (int r = 0; r < a.action_repeat; ++r) {
      float rr;
      if (MATH == MBPO_MATH_REFERENCE) pendulum_step_ref(pc, c, s, w, u, rr);
      else pendulum_step_theta(pc, th, w, u, rr);
      rew = __fadd_rn(rew, rr);
    }
    if (MATH != MBPO_MATH_REFERENCE) sincos_bounded(th, s, c);
    steps = __fadd_rn(steps, rep);
    const bool over = steps >= ep_len;
    const float trunc = over ? (1.0f - done) : 0.0f;
    done = over ? 1.0f : done;
    if (over) { c = f_c; s = f_s; w = f_w; th = f_th; }
    warp_store3(tile, p_nxt, lane, n_valid, c, s, w);
    p_nxt += 3 * E;
    if (live) {
      *p_act = u;
      *p_rew = rew;
      *p_dis = 1.0f - done;
      *p_tru = trunc;
    }
    p_act += E; p_rew += E; p_dis += E; p_tru += E;
  }
  if (live) {
    a.obs[3 * e] = c; a.obs[3 * e + 1] = s; a.obs[3 * e + 2] = w;
    a.steps[e] = steps;
    a.done[e] = done;
  }
  if (e == 0 && a.key_out) { a.key_out[0] = key.k0; a.key_out[1] = key.k1; }
}

}  // namespace mbpo
