// Instantiates the horizon-specialised kernels for H = MBPO_INST_H (see plan_dispatch.h).
#ifndef MBPO_INST_H
#error "compile with -DMBPO_INST_H=<horizon>"
#endif
#include "plan_dispatch.h"

namespace mbpo {

namespace {

constexpr int kH = MBPO_INST_H;
constexpr int kPlanThreads = 256;
// Resident CTAs per SM the register budget is tuned for (shared memory may allow fewer).
constexpr int kMinBlocks = kH <= 20 ? 4 : (kH <= 30 ? 3 : 1);

template <typename Kernel>
int prepare_plan_kernel(Kernel kernel, size_t smem, int B, int* grid_out) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return fail(MBPO_ECUDA, "cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPlanThreads, smem);
  if (e != cudaSuccess) return fail(MBPO_ECUDA, "occupancy query: %s", cudaGetErrorString(e));
  if (per_sm < 1) return fail(MBPO_EUNSUPPORTED, "plan kernel does not fit one SM (smem=%zu)", smem);
  const long long resident = static_cast<long long>(per_sm) * device_sm_count();
  *grid_out = static_cast<int>(B < resident ? B : resident);
  return MBPO_OK;
}

template <int PRNG, int MATH>
int launch(const PlanArgs& a, const MpcArgs* mpc, cudaStream_t st) {
  const size_t smem = PlanSmem<kH>::bytes(a.N, a.Np, a.K);
  int grid = 0;
  if (mpc == nullptr) {
    auto kernel = icem_plan_pendulum_kernel<kH, PRNG, MATH, kPlanThreads, kMinBlocks>;
    const int rc = prepare_plan_kernel(kernel, smem, a.B, &grid);
    if (rc != MBPO_OK) return rc;
    PlanArgs b = a;
    b.zero_value_precomputed = 1;   // best_value_out doubles as the hand-over buffer
    zero_row_value_kernel<MATH><<<(a.B + 127) / 128, 128, 0, st>>>(a.sys, kH, a.P, a.summarize, a.x0, a.B,
                                                                  a.best_value_out);
    const int rc2 = check_launch("zero_row_value_kernel");
    if (rc2 != MBPO_OK) return rc2;
    kernel<<<grid, kPlanThreads, smem, st>>>(b);
    return check_launch("icem_plan_pendulum_kernel");
  }
  auto kernel = icem_mpc_pendulum_kernel<kH, PRNG, MATH, kPlanThreads, kMinBlocks>;
  const int rc = prepare_plan_kernel(kernel, smem, a.B, &grid);
  if (rc != MBPO_OK) return rc;
  kernel<<<grid, kPlanThreads, smem, st>>>(a, *mpc);
  return check_launch("icem_mpc_pendulum_kernel");
}

}  // namespace

template <>
int plan_entry<MBPO_INST_H>(int prng_mode, int math_mode, const PlanArgs& a, const MpcArgs* mpc, cudaStream_t st) {
  switch (prng_mode * 2 + math_mode) {
    case 0: return launch<0, 0>(a, mpc, st);
    case 1: return launch<0, 1>(a, mpc, st);
    case 2: return launch<1, 0>(a, mpc, st);
    default: return launch<1, 1>(a, mpc, st);
  }
}

template <>
int noise_entry<MBPO_INST_H>(int prng_mode, const ScaleTable& tbl, const uint32_t* keys, int M, float* noise_out,
                             uint32_t* bits_out, cudaStream_t st) {
  const int threads = STAGED_THREADS;
  const unsigned blocks = static_cast<unsigned>((M + threads - 1) / threads);
  if (prng_mode == 0) powerlaw_noise_kernel<kH, 0><<<blocks, threads, 0, st>>>(tbl, keys, M, noise_out, bits_out);
  else powerlaw_noise_kernel<kH, 1><<<blocks, threads, 0, st>>>(tbl, keys, M, noise_out, bits_out);
  return check_launch("powerlaw_noise_kernel");
}

template <>
int sample_entry<MBPO_INST_H>(int prng_mode, const ScaleTable& tbl, const uint32_t* carry_key, const float* mean,
                              const float* std_, int N, int Np, int A, float u_min, float u_max, int B,
                              float* actions, uint32_t* next_key, uint32_t* particle_keys, cudaStream_t st) {
  const int threads = STAGED_THREADS;
  const dim3 grid(static_cast<unsigned>(((N + Np) * A + threads - 1) / threads), static_cast<unsigned>(B));
  if (prng_mode == 0)
    sample_actions_kernel<kH, 0><<<grid, threads, 0, st>>>(tbl, carry_key, mean, std_, N, Np, A, u_min, u_max,
                                                           actions, next_key, particle_keys);
  else
    sample_actions_kernel<kH, 1><<<grid, threads, 0, st>>>(tbl, carry_key, mean, std_, N, Np, A, u_min, u_max,
                                                           actions, next_key, particle_keys);
  return check_launch("sample_actions_kernel");
}

}  // namespace mbpo
