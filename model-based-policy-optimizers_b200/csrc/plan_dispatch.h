// Per-horizon kernel entry points.  Each horizon H in MBPO_FOR_EACH_H is compiled in its own
// translation unit (plan_inst.cu with -DMBPO_INST_H=H) so the build parallelises; this header
// declares what those units define.
#pragma once
#include "host_util.h"
#include "icem_kernels.cuh"

namespace mbpo {

// Fused plan (mpc == nullptr) or closed-loop MPC (mpc != nullptr); selects PRNG / MATH
// template variants from prng_mode / math_mode.
template <int H>
int plan_entry(int prng_mode, int math_mode, const PlanArgs& a, const MpcArgs* mpc, cudaStream_t st);

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
