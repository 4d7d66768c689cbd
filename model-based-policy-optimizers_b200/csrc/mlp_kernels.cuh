// Learned MLP-ensemble dynamics forward (BASELINE config 4; template network_utils.py:5-17).
//
//   h0 = swish(inp @ w_in[e] + b_in[e])            float32   (K = x_dim + u_dim = 4)
//   h1 = swish(bf16(h0) @ bf16(w_h[e,0]) + b_h[e,0])   bf16 operands, float32 accumulate
//   h2 = swish(bf16(h1) @ bf16(w_h[e,1]) + b_h[e,1])   bf16 operands, float32 accumulate
//   delta = h2 @ w_out[e] + b_out[e]               float32   (N = x_dim = 3)
//
// Round-1 bring-up kernel: CUDA-core evaluation used to establish parity of the data layout
// and rounding contract against the oracle.  The tcgen05/TMA kernel replaces the two hidden
// GEMMs (stage 4).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/mbpo_b200.h"

namespace mbpo {

__device__ __forceinline__ float swishf(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

constexpr int MLP_ROWS = 16;     // rows per CTA
constexpr int MLP_THREADS = 256; // one thread per hidden unit

// One CTA evaluates MLP_ROWS consecutive rows; thread j owns hidden unit j.
__global__ void __launch_bounds__(MLP_THREADS) mlp_forward_simt_kernel(const MbpoMlpEnsembleParams p,
                                                                       const float* __restrict__ inp,
                                                                       const int32_t* __restrict__ member, int R,
                                                                       float* __restrict__ delta_out) {
  extern __shared__ float mlp_sm[];
  const int Hd = p.hidden;
  const int In = p.x_dim + p.u_dim;
  float* act_a = mlp_sm;                // [MLP_ROWS][Hd]
  float* act_b = mlp_sm + MLP_ROWS * Hd; // [MLP_ROWS][Hd]
  const int row0 = blockIdx.x * MLP_ROWS;
  const int j = threadIdx.x;
  for (int r = 0; r < MLP_ROWS; ++r) {
    const int row = row0 + r;
    if (row >= R) break;
    const int e = member[row];
    if (j < Hd) {
      float acc = p.b_in[e * Hd + j];
      for (int k = 0; k < In; ++k) acc = fmaf(inp[static_cast<size_t>(row) * In + k], p.w_in[(e * In + k) * Hd + j], acc);
      act_a[r * Hd + j] = bf16_round(swishf(acc));
    }
  }
  __syncthreads();
  float* src = act_a;
  float* dst = act_b;
  for (int l = 0; l < 2; ++l) {
    for (int r = 0; r < MLP_ROWS; ++r) {
      const int row = row0 + r;
      if (row >= R) break;
      const int e = member[row];
      if (j < Hd) {
        const __nv_bfloat16* w =
            reinterpret_cast<const __nv_bfloat16*>(p.w_h) + ((static_cast<size_t>(e) * 2 + l) * Hd + j) * Hd;
        float acc = 0.0f;
        for (int k = 0; k < Hd; ++k) acc = fmaf(src[r * Hd + k], __bfloat162float(w[k]), acc);
        acc += p.b_h[(e * 2 + l) * Hd + j];
        const float a = swishf(acc);
        dst[r * Hd + j] = (l == 0) ? bf16_round(a) : a;
      }
    }
    __syncthreads();
    float* t = src; src = dst; dst = t;
  }
  // output layer: thread (r, o)
  const int X = p.x_dim;
  for (int idx = j; idx < MLP_ROWS * X; idx += MLP_THREADS) {
    const int r = idx / X, o = idx % X;
    const int row = row0 + r;
    if (row >= R) continue;
    const int e = member[row];
    float acc = p.b_out[e * X + o];
    for (int k = 0; k < Hd; ++k) acc = fmaf(src[r * Hd + k], p.w_out[(static_cast<size_t>(e) * Hd + k) * X + o], acc);
    delta_out[static_cast<size_t>(row) * X + o] = acc;
  }
}

inline int launch_mlp_forward(const MbpoMlpEnsembleParams& p, const float* inp, const int32_t* member, int R,
                              float* delta_out, cudaStream_t st, char* err, size_t errlen) {
  if (p.hidden < 1 || p.hidden > MLP_THREADS || p.x_dim < 1 || p.u_dim < 0 || p.num_members < 1) {
    snprintf(err, errlen, "mlp_dynamics_forward: unsupported shape (hidden=%d, x_dim=%d, u_dim=%d, members=%d)",
             p.hidden, p.x_dim, p.u_dim, p.num_members);
    return MBPO_EUNSUPPORTED;
  }
  const size_t smem = static_cast<size_t>(2) * MLP_ROWS * p.hidden * sizeof(float);
  const unsigned blocks = static_cast<unsigned>((R + MLP_ROWS - 1) / MLP_ROWS);
  mlp_forward_simt_kernel<<<blocks, MLP_THREADS, smem, st>>>(p, inp, member, R, delta_out);
  return MBPO_OK;
}

}  // namespace mbpo
