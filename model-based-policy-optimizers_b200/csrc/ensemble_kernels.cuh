// Stage 4, fused: open-loop rollouts through a learned MLP-ensemble System (BASELINE config 4).
//
// vmap(vmap(rollout_actions)) (mbpo/utils/optimizer_utils.py:11-59) + the particle summary of
// iCemTO.objective (icem_optimizer.py:155-160) for a System whose step is
//     x_next = x + MLP_e([x, u])           (template: mbpo/utils/network_utils.py:5-17, swish)
//     reward = pendulum reward on (x, u)   (rewards/pendulum_reward.py:27-42)
// with particle p rolled through ensemble member p (num_particles == num_members).
//
// A CTA PAIR (thread-block cluster of 2, tcgen05 cta_group::2) owns 256 candidate rows; each
// CTA owns 128 rows = its 128 TMEM lanes, four threads per row (one per quarter of the hidden units).  For every member e the pair keeps BOTH
// hidden-layer weight matrices resident in shared memory for the whole horizon, split along N
// across the pair (each CTA holds W[e,l][128 rows of N, 256 K] = 64 KB per layer, fetched by TMA as
// 32 K-chunk slabs), so the only per-step traffic is the action stream.  Per horizon step:
//   layer 0 (K = 4) on CUDA cores -> bf16 activations into the canonical K-major A tile in smem
//   16 x tcgen05.mma.cta_group::2 (M256 N256 K16), issued by one thread of the leader CTA;
//        completion multicast to both CTAs' mbarriers by tcgen05.commit
//   epilogue 1: tcgen05.ld accumulator rows, bias + swish, bf16 -> A tile (next layer's operand)
//   16 x tcgen05.mma with W[e,1]
//   epilogue 2: bias + swish, output layer (N = 3) on CUDA cores, x += delta, reward accumulated
// The state never leaves registers; returns are averaged (or maxed) over members in registers in
// member order, so the result is deterministic and needs no workspace.
#pragma once
#include "mathx.cuh"
#include "mlp_tc_kernels.cuh"
#include "pendulum.cuh"

namespace mbpo {
namespace ens {

using namespace tc;

constexpr int ENS_THREADS = 512;              // four threads per row: warps 4q..4q+3 own accumulator columns [64q, 64q+64)
constexpr int ENS_SPLIT = ENS_THREADS / 128;  // threads per row
constexpr uint32_t WH_BYTES = 128 * HID * 2;  // one layer's N-half: 65536
constexpr uint32_t WH_LBO = 128 * 16;         // 2048: next K-chunk of a weight half

struct Smem {
  static constexpr uint32_t A = 0;                        // 128 x 256 bf16 activations
  static constexpr uint32_t W1 = A + A_BYTES;             // W[e,0] rows [128*rank, +128)
  static constexpr uint32_t W2 = W1 + WH_BYTES;           // W[e,1] rows [128*rank, +128)
  static constexpr uint32_t W_IN = W2 + WH_BYTES;         // float [4][256]
  static constexpr uint32_t B_IN = W_IN + 4 * HID * 4;    // float [256]
  static constexpr uint32_t B_H = B_IN + HID * 4;         // float [2][256]
  static constexpr uint32_t W_OUT = B_H + 2 * HID * 4;    // float [256][4]
  static constexpr uint32_t B_OUT = W_OUT + HID * 4 * 4;  // float [4]
  static constexpr uint32_t DELTA = B_OUT + 16;           // float4 [ENS_SPLIT][128]: partial output-layer sums
  static constexpr uint32_t BARS = DELTA + ENS_SPLIT * TILE_M * 16;  // bar_w, bar_mma, bar_a
  static constexpr uint32_t TMEM_PTR = BARS + 32;
  static constexpr uint32_t TOTAL = TMEM_PTR + 16;
};
static_assert(Smem::TOTAL <= 227 * 1024, "ensemble rollout shared memory plan exceeds 227 KB");

struct EnsArgs {
  int num_members, R, H, M, summarize;  // R = B * M rows; row r belongs to problem r / M
  const float* w_in;    // [E, 4, 256]
  const float* b_in;    // [E, 256]
  const float* b_h;     // [E, 2, 256]
  const float* w_out;   // [E, 256, 3]
  const float* b_out;   // [E, 3]
  const float* x0;      // [B, 3]
  const float* actions; // [R, H]
  float* returns_out;   // [R]
  MbpoPendulumParams reward;
};

// swish(z) = z * sigmoid(z) with sigmoid(z) = 0.5 + 0.5 * tanh(z / 2): one MUFU (tanh.approx) instead of
// two (ex2 + rcp); the activation pipeline is MUFU-throughput bound.  |error| <= ~2.5e-4 * |z|, well
// inside the bf16 rounding (2^-9 relative) the activations receive next.
__device__ __forceinline__ float swish_tanh(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return z * fmaf(t, 0.5f, 0.5f);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ENS_THREADS, 1)
    ensemble_rollout_kernel(const __grid_constant__ EnsArgs a, const __grid_constant__ CUtensorMap w_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int lrow = tid & (TILE_M - 1);   // row within the CTA's tile = TMEM lane
  const int half = tid >> 7;             // which 256 / ENS_SPLIT hidden units / accumulator columns this thread owns
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  float* s_w_in = reinterpret_cast<float*>(smem + Smem::W_IN);
  float* s_b_in = reinterpret_cast<float*>(smem + Smem::B_IN);
  float* s_b_h = reinterpret_cast<float*>(smem + Smem::B_H);
  float* s_w_out = reinterpret_cast<float*>(smem + Smem::W_OUT);
  float* s_b_out = reinterpret_cast<float*>(smem + Smem::B_OUT);
  float4* s_delta = reinterpret_cast<float4*>(smem + Smem::DELTA);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + Smem::BARS);  // weights landed (this CTA)
  uint64_t* bar_mma = bar_w + 1;                                     // accumulator complete (multicast commit)
  uint64_t* bar_a = bar_w + 2;                                       // leader only: both CTAs' A tiles written
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);

  // ---- one-time setup --------------------------------------------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_a, 2);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + half * (HID / ENS_SPLIT);
  const uint32_t a_addr = smem_u32(smem + Smem::A);
  const uint32_t w_addr[2] = {smem_u32(smem + Smem::W1), smem_u32(smem + Smem::W2)};
  const uint32_t bar_a_leader = map_to_cta(smem_u32(bar_a), 0);
  constexpr uint32_t IDESC = umma_idesc_bf16(2 * TILE_M, HID);
  uint32_t phase_w = 0, phase_mma = 0, phase_a = 0;
  const PendulumConsts pc(a.reward);
  const float inv_h = 1.0f;  // (mean over the horizon taken with __fdiv_rn below)
  (void)inv_h;

  const int num_groups = (a.R + 2 * TILE_M - 1) / (2 * TILE_M);
  for (int group = blockIdx.x >> 1; group < num_groups; group += gridDim.x >> 1) {
    const int row = group * 2 * TILE_M + static_cast<int>(rank) * TILE_M + lrow;
    const bool valid = row < a.R;
    const int rr = valid ? row : a.R - 1;
    const int b = rr / a.M;
    const float x_init[3] = {a.x0[3 * b], a.x0[3 * b + 1], a.x0[3 * b + 2]};
    const float* act = a.actions + static_cast<size_t>(rr) * a.H;
    float summary = 0.0f;

    for (int e = 0; e < a.num_members; ++e) {
      // ---- member e: both weight halves by TMA, small fp32 parameters by the threads ------------
      __syncthreads();  // previous member's parameters / A tile no longer in use by this CTA
      if (tid == 0) {
        mbar_expect_tx(bar_w, 2 * WH_BYTES);
        for (int l = 0; l < 2; ++l)
          for (int kc = 0; kc < KCHUNKS; ++kc)
            tma_load_2d(smem + (l ? Smem::W2 : Smem::W1) + kc * WH_LBO, &w_map, kc * 8,
                        (e * 2 + l) * HID + static_cast<int>(rank) * 128, bar_w);
      }
      for (int i = tid; i < 4 * HID; i += ENS_THREADS) s_w_in[i] = a.w_in[static_cast<size_t>(e) * 4 * HID + i];
      for (int i = tid; i < HID; i += ENS_THREADS) s_b_in[i] = a.b_in[e * HID + i];
      for (int i = tid; i < 2 * HID; i += ENS_THREADS) s_b_h[i] = a.b_h[e * 2 * HID + i];
      for (int i = tid; i < HID * 3; i += ENS_THREADS) s_w_out[(i / 3) * 4 + (i % 3)] = a.w_out[static_cast<size_t>(e) * HID * 3 + i];
      if (tid < 3) s_b_out[tid] = a.b_out[e * 3 + tid];
      mbar_wait(bar_w, phase_w);  // every thread observes the TMA completion (async-proxy writes visible)
      phase_w ^= 1;
      __syncthreads();

      float x[3] = {x_init[0], x_init[1], x_init[2]};
      float acc = 0.0f;
#pragma unroll 1
      for (int t = 0; t < a.H; ++t) {
        const float u = __ldg(act + t);
        // reward on the current state and the raw action (pendulum_reward.py:32-40)
        acc = __fadd_rn(acc, reward_from(pc, atan2_bounded(x[1], x[0]), x[2], u));
        // ---- layer 0 on CUDA cores ---------------------------------------------------------------
#pragma unroll 2
        for (int kc = half * (KCHUNKS / ENS_SPLIT); kc < (half + 1) * (KCHUNKS / ENS_SPLIT); ++kc) {
          float h[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int n = kc * 8 + j;
            float v = s_b_in[n];
            v = fmaf(x[0], s_w_in[n], v);
            v = fmaf(x[1], s_w_in[HID + n], v);
            v = fmaf(x[2], s_w_in[2 * HID + n], v);
            v = fmaf(u, s_w_in[3 * HID + n], v);
            h[j] = swish_tanh(v);
          }
          uint4 pk;
          pk.x = pack_bf16(h[0], h[1]); pk.y = pack_bf16(h[2], h[3]);
          pk.z = pack_bf16(h[4], h[5]); pk.w = pack_bf16(h[6], h[7]);
          *reinterpret_cast<uint4*>(smem + Smem::A + kc * A_LBO + lrow * 16) = pk;
        }
        float delta[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          // ---- publish this CTA's A tile to the tensor core, tell the leader -----------------------
          tc_fence_before();
          fence_proxy_async();
          __syncthreads();
          if (tid == 0) mbar_arrive_cluster(bar_a_leader);
          if (leader && tid == 0) {
            mbar_wait_cluster(bar_a, phase_a);
            tc_fence_after();
#pragma unroll
            for (int s = 0; s < HID / 16; ++s) {
              const uint64_t da = umma_desc(a_addr + s * 2 * A_LBO, A_LBO, SBO);
              const uint64_t db = umma_desc(w_addr[layer] + s * 2 * WH_LBO, WH_LBO, SBO);
              umma_bf16_ss_2sm(tmem_base, da, db, IDESC, s > 0 ? 1u : 0u);
            }
            umma_commit_2sm(bar_mma);
          }
          phase_a ^= 1;
          mbar_wait(bar_mma, phase_mma);
          phase_mma ^= 1;
          tc_fence_after();
          // ---- epilogue ---------------------------------------------------------------------------------
          const float* bias = s_b_h + layer * HID;
#pragma unroll 1
          for (int c = 0; c < HID / 32 / ENS_SPLIT; ++c) {
            const int cg = half * (HID / 32 / ENS_SPLIT) + c;        // global 32-column chunk
            uint32_t v[32];
            tmem_ld32(tmem_row + c * 32, v);
            float h[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) h[j] = swish_tanh(__uint_as_float(v[j]) + bias[cg * 32 + j]);
            if (layer == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pk;
                pk.x = pack_bf16(h[q * 8 + 0], h[q * 8 + 1]); pk.y = pack_bf16(h[q * 8 + 2], h[q * 8 + 3]);
                pk.z = pack_bf16(h[q * 8 + 4], h[q * 8 + 5]); pk.w = pack_bf16(h[q * 8 + 6], h[q * 8 + 7]);
                *reinterpret_cast<uint4*>(smem + Smem::A + (cg * 4 + q) * A_LBO + lrow * 16) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float4 wo = *reinterpret_cast<const float4*>(s_w_out + (cg * 32 + j) * 4);
                delta[0] = fmaf(h[j], wo.x, delta[0]);
                delta[1] = fmaf(h[j], wo.y, delta[1]);
                delta[2] = fmaf(h[j], wo.z, delta[2]);
              }
            }
          }
        }
        // the next step's layer 0 overwrites A only after this CTA's MMA reads completed (bar_mma) and
        // tcgen05.ld of D completed (tcgen05.wait::ld inside tmem_ld32); D is rewritten only after the
        // next "A ready" handshake, which every thread precedes with tcgen05.fence::before_thread_sync.
        // combine the column quarters of the output layer (fixed order: q0 + q1 + ... + bias)
        s_delta[half * TILE_M + lrow] = make_float4(delta[0], delta[1], delta[2], 0.0f);
        __syncthreads();
        float4 d = s_delta[lrow];
#pragma unroll
        for (int q = 1; q < ENS_SPLIT; ++q) {
          const float4 dq = s_delta[q * TILE_M + lrow];
          d.x += dq.x; d.y += dq.y; d.z += dq.z;
        }
        x[0] = __fadd_rn(x[0], d.x + s_b_out[0]);
        x[1] = __fadd_rn(x[1], d.y + s_b_out[1]);
        x[2] = __fadd_rn(x[2], d.z + s_b_out[2]);
      }
      const float ret = __fdiv_rn(acc, static_cast<float>(a.H));
      if (a.summarize == MBPO_SUMMARIZE_MAX) summary = (e == 0) ? ret : fmaxf(summary, ret);
      else summary = __fadd_rn(summary, ret);
    }
    if (a.summarize != MBPO_SUMMARIZE_MAX) summary = __fdiv_rn(summary, static_cast<float>(a.num_members));
    if (valid && half == 0) a.returns_out[row] = summary;
  }

  // ---- teardown --------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

inline int launch_ensemble_rollout(const MbpoMlpEnsembleParams& p, int horizon, const float* x0, const float* actions,
                                   int B, int M, int summarize, float* returns_out, cudaStream_t st, char* err,
                                   size_t errlen) {
  if (p.hidden != HID || p.x_dim != 3 || p.u_dim != 1 || p.num_members < 1) {
    snprintf(err, errlen,
             "ensemble rollout: the tcgen05 kernel needs hidden == 256, x_dim == 3, u_dim == 1 (got %d, %d, %d)",
             p.hidden, p.x_dim, p.u_dim);
    return MBPO_EUNSUPPORTED;
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    snprintf(err, errlen, "ensemble rollout: cuTensorMapEncodeTiled is not available from the driver");
    return MBPO_ECUDA;
  }
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(HID), static_cast<cuuint64_t>(p.num_members) * 2 * HID};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(HID) * 2};
  const cuuint32_t box[2] = {8, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(p.w_h), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(err, errlen, "ensemble rollout: cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(cr));
    return MBPO_ECUDA;
  }
  EnsArgs a;
  a.num_members = p.num_members; a.R = B * M; a.H = horizon; a.M = M; a.summarize = summarize;
  a.w_in = p.w_in; a.b_in = p.b_in; a.b_h = p.b_h; a.w_out = p.w_out; a.b_out = p.b_out;
  a.x0 = x0; a.actions = actions; a.returns_out = returns_out; a.reward = p.reward;
  cudaError_t ce = cudaFuncSetAttribute(ensemble_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(Smem::TOTAL));
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "ensemble rollout: smem attribute: %s", cudaGetErrorString(ce));
    return MBPO_ECUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int groups = (a.R + 2 * TILE_M - 1) / (2 * TILE_M);
  const int pairs = groups < sms / 2 ? groups : sms / 2;
  ensemble_rollout_kernel<<<2 * pairs, ENS_THREADS, Smem::TOTAL, st>>>(a, map);
  return MBPO_OK;
}

}  // namespace ens
}  // namespace mbpo
