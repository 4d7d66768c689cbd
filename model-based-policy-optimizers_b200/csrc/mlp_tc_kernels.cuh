// Stage 4: learned MLP-ensemble dynamics forward on the 5th-generation tensor cores.
//
//   h0 = swish(inp @ w_in[e] + b_in[e])                 CUDA cores, float32   (K = x_dim + u_dim = 4)
//   h1 = swish(bf16(h0) @ bf16(w_h[e,0]) + b_h[e,0])    tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM)
//   h2 = swish(bf16(h1) @ bf16(w_h[e,1]) + b_h[e,1])    tcgen05.mma
//   delta = bf16(h2) @ w_out[e] + b_out[e]              CUDA cores, float32   (N = x_dim = 3)
//
// One CTA (128 threads = 128 TMEM lanes) owns a tile of 128 rows; thread r owns row r.
//   * A operand (activations, 128 x 256 bf16) lives in shared memory in the canonical K-major
//     no-swizzle UMMA layout, written by the epilogue itself: 16-byte K-chunks, chunk kc of row r at
//     kc * 2048 + r * 16  (core matrix = 8 rows x 16 B contiguous; SBO = 128 B, LBO = 2048 B).
//     A warp's 32 rows write 512 contiguous bytes per chunk: conflict-free 128-bit stores.
//   * B operand (weights, 256(out) x 256(in) bf16, K-major in HBM) is fetched by TMA: a 2-D
//     tensor map with box {8 elements, 256 rows} lands each K-chunk as a 4 KB slab at kc * 4096 --
//     the same canonical layout (SBO = 128 B, LBO = 4096 B) with no shuffling by any thread.
//   * D (128 x 256 fp32) is 256 TMEM columns; 16 tcgen05.mma (M128 N256 K16) per layer are issued
//     by one thread, completion is tracked with tcgen05.commit -> mbarrier, the epilogue reads the
//     accumulator with tcgen05.ld.32x32b (each thread: its own row, 32 columns at a time).
// Rows of a tile may belong to different ensemble members: the tile is then evaluated once per
// member present (a homogeneous tile -- the iCEM layout -- takes one pass).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/mbpo_b200.h"

namespace mbpo {
namespace tc {

constexpr int TILE_M = 128;
constexpr int HID = 256;
constexpr int KCHUNKS = HID / 8;                 // 16-byte chunks along K
constexpr uint32_t A_BYTES = TILE_M * HID * 2;   // 65536
constexpr uint32_t W_BYTES = HID * HID * 2;      // 131072
constexpr uint32_t A_LBO = TILE_M * 16;          // 2048: next K-chunk of A
constexpr uint32_t W_LBO = HID * 16;             // 4096: next K-chunk of W
constexpr uint32_t SBO = 128;                    // next 8-row core matrix
constexpr int TMEM_COLS = 256;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// [0,14), leading byte offset [16,30), stride byte offset [32,46) -- all in 16-byte units --,
// version 1 at [46,48), layout type 0 (SWIZZLE_NONE) at [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16: D = F32 (1 @ [4,6)), A = B = BF16
// (1 @ [7,10), [10,13)), both K-major (0 @ 15, 16), N >> 3 @ [17,23), M >> 4 @ [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float swish_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&p);
}

// ---- shared-memory plan ------------------------------------------------------------------------
struct Smem {
  static constexpr uint32_t A = 0;                       // activations, canonical K-major
  static constexpr uint32_t W = A + A_BYTES;             // one 256 x 256 weight matrix
  static constexpr uint32_t W_IN = W + W_BYTES;          // float [4][256]  (x_dim + u_dim <= 4 rows used)
  static constexpr uint32_t B_IN = W_IN + 8 * HID * 4;   // float [256]   (room for up to 8 input features)
  static constexpr uint32_t B_H = B_IN + HID * 4;        // float [2][256]
  static constexpr uint32_t W_OUT = B_H + 2 * HID * 4;   // float [256][4] (x_dim <= 4, padded)
  static constexpr uint32_t B_OUT = W_OUT + HID * 4 * 4; // float [4]
  static constexpr uint32_t BARS = B_OUT + 16;           // 2 mbarriers
  static constexpr uint32_t TMEM_PTR = BARS + 16;
  static constexpr uint32_t TOTAL = TMEM_PTR + 16;
};
static_assert(Smem::TOTAL <= 227 * 1024, "stage-4 shared memory plan exceeds 227 KB");

struct MlpArgs {
  int num_members, x_dim, u_dim, R;
  const float* w_in;   // [E, In, 256]
  const float* b_in;   // [E, 256]
  const float* b_h;    // [E, 2, 256]
  const float* w_out;  // [E, 256, X]
  const float* b_out;  // [E, X]
  const float* inp;    // [R, In]
  const int32_t* member;  // [R]
  float* delta_out;    // [R, X]
};

__global__ void __launch_bounds__(TILE_M, 1)
    mlp_forward_tc_kernel(const __grid_constant__ MlpArgs a, const __grid_constant__ CUtensorMap w_map) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  float* s_w_in = reinterpret_cast<float*>(smem + Smem::W_IN);
  float* s_b_in = reinterpret_cast<float*>(smem + Smem::B_IN);
  float* s_b_h = reinterpret_cast<float*>(smem + Smem::B_H);
  float* s_w_out = reinterpret_cast<float*>(smem + Smem::W_OUT);
  float* s_b_out = reinterpret_cast<float*>(smem + Smem::B_OUT);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + Smem::BARS);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_PTR);
  const int In = a.x_dim + a.u_dim, X = a.x_dim;

  // ---- one-time setup: TMEM allocation (warp 0), mbarriers (thread 0) ---------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);  // this warp's 32 lanes
  const uint32_t a_addr = smem_u32(smem + Smem::A), w_addr = smem_u32(smem + Smem::W);
  constexpr uint32_t IDESC = umma_idesc_bf16(TILE_M, HID);
  uint32_t phase_w = 0, phase_mma = 0;

  const int num_tiles = (a.R + TILE_M - 1) / TILE_M;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int row = tile * TILE_M + tid;
    const bool valid = row < a.R;
    const int e_row = valid ? a.member[row] : -1;
    float xin[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) xin[k] = (valid && k < In) ? a.inp[static_cast<size_t>(row) * In + k] : 0.0f;

    for (int e = 0; e < a.num_members; ++e) {
      if (!__syncthreads_or(e_row == e)) continue;  // CTA-uniform; also orders reuse of the buffers below

      // ---- member e: TMA for W[e,0]; small fp32 parameters to shared memory ----------------------
      if (tid == 0) {
        mbar_expect_tx(bar_w, W_BYTES);
        for (int kc = 0; kc < KCHUNKS; ++kc)
          tma_load_2d(smem + Smem::W + kc * W_LBO, &w_map, kc * 8, (e * 2 + 0) * HID, bar_w);
      }
      for (int i = tid; i < In * HID; i += TILE_M) s_w_in[i] = a.w_in[static_cast<size_t>(e) * In * HID + i];
      for (int i = tid; i < HID; i += TILE_M) s_b_in[i] = a.b_in[e * HID + i];
      for (int i = tid; i < 2 * HID; i += TILE_M) s_b_h[i] = a.b_h[e * 2 * HID + i];
      for (int i = tid; i < HID * X; i += TILE_M)
        s_w_out[(i / X) * 4 + (i % X)] = a.w_out[static_cast<size_t>(e) * HID * X + i];
      if (tid < X) s_b_out[tid] = a.b_out[e * X + tid];
      __syncthreads();

      // ---- layer 0 on CUDA cores: thread = row, 8 hidden units per 16-byte chunk ------------------
#pragma unroll 2
      for (int kc = 0; kc < KCHUNKS; ++kc) {
        float h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = s_b_in[kc * 8 + j];
        for (int k = 0; k < In; ++k) {
          const float xk = xin[k];
#pragma unroll
          for (int j = 0; j < 8; ++j) h[j] = fmaf(xk, s_w_in[k * HID + kc * 8 + j], h[j]);
        }
        uint4 pk;
        pk.x = pack_bf16(swish_f(h[0]), swish_f(h[1]));
        pk.y = pack_bf16(swish_f(h[2]), swish_f(h[3]));
        pk.z = pack_bf16(swish_f(h[4]), swish_f(h[5]));
        pk.w = pack_bf16(swish_f(h[6]), swish_f(h[7]));
        *reinterpret_cast<uint4*>(smem + Smem::A + kc * A_LBO + tid * 16) = pk;
      }
      fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor core (async proxy)
      __syncthreads();

      float out[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 1
      for (int layer = 0; layer < 2; ++layer) {
        // ---- 16 x tcgen05.mma (M128 N256 K16), one issuing thread ----------------------------------
        if (tid == 0) {
          mbar_wait(bar_w, phase_w);  // this layer's weights have landed
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < HID / 16; ++s) {
            const uint64_t da = umma_desc(a_addr + s * 2 * A_LBO, A_LBO, SBO);
            const uint64_t db = umma_desc(w_addr + s * 2 * W_LBO, W_LBO, SBO);
            umma_bf16_ss(tmem_base, da, db, IDESC, s > 0 ? 1u : 0u);
          }
          umma_commit(bar_mma);  // implies tcgen05.fence::before_thread_sync
        }
        phase_w ^= 1;
        mbar_wait(bar_mma, phase_mma);  // accumulator complete; A and W are free again
        phase_mma ^= 1;
        tc_fence_after();
        if (layer == 0 && tid == 0) {    // prefetch W[e,1] under the epilogue
          mbar_expect_tx(bar_w, W_BYTES);
          for (int kc = 0; kc < KCHUNKS; ++kc)
            tma_load_2d(smem + Smem::W + kc * W_LBO, &w_map, kc * 8, (e * 2 + 1) * HID, bar_w);
        }
        // ---- epilogue: TMEM -> registers, bias + swish ------------------------------------------------
        const float* bias = s_b_h + layer * HID;
#pragma unroll 1
        for (int c = 0; c < HID / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tmem_row + c * 32, v);
          float h[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) h[j] = swish_f(__uint_as_float(v[j]) + bias[c * 32 + j]);
          if (layer == 0) {  // next layer's A operand, bf16
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 pk;
              pk.x = pack_bf16(h[q * 8 + 0], h[q * 8 + 1]);
              pk.y = pack_bf16(h[q * 8 + 2], h[q * 8 + 3]);
              pk.z = pack_bf16(h[q * 8 + 4], h[q * 8 + 5]);
              pk.w = pack_bf16(h[q * 8 + 6], h[q * 8 + 7]);
              *reinterpret_cast<uint4*>(smem + Smem::A + (c * 4 + q) * A_LBO + tid * 16) = pk;
            }
          } else {           // output layer on CUDA cores, float32
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float4 wo = *reinterpret_cast<const float4*>(s_w_out + (c * 32 + j) * 4);
              // the activations entering the output layer are bf16 too (the fused rollout kernel runs
              // that layer on the tensor cores as well); the weights stay float32
              const float hj = __bfloat162float(__float2bfloat16_rn(h[j]));
              out[0] = fmaf(hj, wo.x, out[0]);
              out[1] = fmaf(hj, wo.y, out[1]);
              out[2] = fmaf(hj, wo.z, out[2]);
              out[3] = fmaf(hj, wo.w, out[3]);
            }
          }
        }
        tc_fence_before();     // order the tcgen05.ld above before the next MMA overwrites the accumulator
        fence_proxy_async();
        __syncthreads();
      }
      if (e_row == e) {
        for (int o = 0; o < X; ++o) a.delta_out[static_cast<size_t>(row) * X + o] = out[o] + s_b_out[o];
      }
    }
  }

  // ---- teardown --------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- host side: tensor map for w_h viewed as a [E*2*256, 256] bf16 matrix ----------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int launch_mlp_forward_tc(const MbpoMlpEnsembleParams& p, const float* inp, const int32_t* member, int R,
                                 float* delta_out, cudaStream_t st, char* err, size_t errlen) {
  if (p.hidden != HID || p.x_dim < 1 || p.x_dim > 4 || p.u_dim < 0 || p.x_dim + p.u_dim > 8 || p.num_members < 1) {
    snprintf(err, errlen,
             "mlp_dynamics_forward: the tcgen05 kernel needs hidden == 256, x_dim <= 4, x_dim + u_dim <= 8 "
             "(got hidden=%d, x_dim=%d, u_dim=%d, members=%d)",
             p.hidden, p.x_dim, p.u_dim, p.num_members);
    return MBPO_EUNSUPPORTED;
  }
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    snprintf(err, errlen, "mlp_dynamics_forward: cuTensorMapEncodeTiled is not available from the driver");
    return MBPO_ECUDA;
  }
  CUtensorMap map;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(HID), static_cast<cuuint64_t>(p.num_members) * 2 * HID};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(HID) * 2};
  const cuuint32_t box[2] = {8, HID};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<uint16_t*>(p.w_h), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    snprintf(err, errlen, "mlp_dynamics_forward: cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(cr));
    return MBPO_ECUDA;
  }
  MlpArgs a;
  a.num_members = p.num_members; a.x_dim = p.x_dim; a.u_dim = p.u_dim; a.R = R;
  a.w_in = p.w_in; a.b_in = p.b_in; a.b_h = p.b_h; a.w_out = p.w_out; a.b_out = p.b_out;
  a.inp = inp; a.member = member; a.delta_out = delta_out;
  cudaError_t ce = cudaFuncSetAttribute(mlp_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(Smem::TOTAL));
  if (ce != cudaSuccess) {
    snprintf(err, errlen, "mlp_dynamics_forward: smem attribute: %s", cudaGetErrorString(ce));
    return MBPO_ECUDA;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (R + TILE_M - 1) / TILE_M;
  const int grid = tiles < sms ? tiles : sms;
  mlp_forward_tc_kernel<<<grid, TILE_M, Smem::TOTAL, st>>>(a, map);
  return MBPO_OK;
}

}  // namespace tc
}  // namespace mbpo
