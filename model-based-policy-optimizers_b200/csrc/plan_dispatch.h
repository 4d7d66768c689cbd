// Per-horizon kernel entry points.  Each horizon H in MBPO_FOR_EACH_H is compiled in its own
// translation unit (plan_inst.cu with -DMBPO_INST_H=H) so the build parallelises; this header
// declares what those units define.
#pragma once
#include "host_util.h"
#include "icem_cluster_kernels.cuh"
#include "icem_kernels.cuh"

namespace mbpo {

// Fused plan (mpc == nullptr) or closed-loop MPC (mpc != nullptr); selects PRNG / MATH
// template variants from prng_mode / math_mode.  cluster > 1: every problem is spread over a thread-block cluster
// of that many CTAs (icem_cluster_kernels.cuh; same bits, for few problems); 0 / 1: one CTA per problem.
template <int H>
int plan_entry(int prng_mode, int math_mode, const PlanArgs& a, const MpcArgs* mpc, cudaStream_t st, int cluster);

// The same for a horizon without an unrolled instance (plan_rt.cu; PlanArgs::H).
int plan_entry_rt(int prng_mode, int math_mode, const PlanArgs& a, const MpcArgs* mpc, cudaStream_t st);

// Fused plan over the general Systems (MBPO_SYSTEM_NOISY_PENDULUM, MBPO_SYSTEM_POINT_MASS): icem_plan_general_kernel.
template <int H>
int general_plan_entry(int system_kind, int prng_mode, const PlanArgs& a, cudaStream_t st);

// Cluster size the library picks for B problems of N candidates (0: one CTA per problem).
inline int plan_cluster_size(int B, int N) {
  const int sms = device_sm_count();
  if (B <= 0 || B >= sms) return 0;
  // one or two problems: the non-portable 16 (32 candidates per CTA at N = 512: one cooperative sampling chunk);
  // otherwise the largest portable size that still gives every cluster its own SMs
  int c = (B <= 2 && (N + 15) / 16 >= 32) ? 16 : 8;
  while (c > 1 && (B * c > sms || (N + c - 1) / c < 32)) c >>= 1;
  if (c <= 1 || (N + c - 1) / c > 256) return 0;
  return c;
}

// vmap(powerlaw_psd_gaussian) over M keys.
template <int H>
int noise_entry(int prng_mode, const ScaleTable& tbl, const uint32_t* keys, int M, float* noise_out,
                uint32_t* bits_out, cudaStream_t st);

// One iteration of key plumbing + sampling for B problems.
template <int H>
int sample_entry(int prng_mode, const ScaleTable& tbl, const uint32_t* carry_key, const float* mean,
                 const float* std_, int N, int Np, int A, float u_min, float u_max, int B, float* actions,
                 uint32_t* next_key, uint32_t* particle_keys, cudaStream_t st);

}  // namespace mbpo
