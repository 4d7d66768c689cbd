// Policy-in-the-loop data collection, latency-optimised: one 128-env tile per CTA, the tile's work spread over
// sixteen producer warps.
//
// actor_rollout_tc_kernel (actor_tc_kernels.cuh) gives every env one thread that walks all 64 columns of every layer:
// right when there are enough envs to put four tiles on every SM, but a launch is T sequential policy steps whatever
// the grid, and one tile alone on an SM steps in 6.6 us (a single warp per scheduler issuing dependent instructions).
// The reference's own configurations are in that regime (tests/test_sac.py: num_envs = 32, 20 steps per collection),
// and so is every shard of config 3 at 4 or 8 GPUs (8,192 envs = 64 tiles for 148 SMs).  Here
//
//   producers     16 warps: warp = 4 * q + lane_quarter; the four threads (q = 0..3, row) share the 64 columns of their
//                 env's row in every layer: layer 0, the bias + swish epilogues and the TF32 hi / lo split run four
//                 wide per env.  TMEM lane quarter = warp % 4, as the hardware requires.
//   A operand     four 16 KB slots.  A stage runs in four time slices: in slice j thread (q, row) produces columns
//                 16 j + 4 q .. + 3 -- one 16-byte chunk of slot j -- so slot j is complete a quarter of the way through
//                 the stage and its MMAs run under the production of slot j + 1; the issuer takes the slots in the
//                 order 0..3 (the accumulation order of the four-tile kernel).  A slot is rewritten only after
//                 `layer_done` of the MMAs that read it, so there are no slot_free barriers.
//   output layer  every thread folds its 16 columns into a partial (loc, scale) pair; the four partials of a row are
//                 added in quarter order by the q = 0 thread -- the four-tile kernel adds its quarters in the same
//                 order, so both kernels (and therefore sharded and unsharded launches) agree bit for bit.
//   head, env     the q = 0 warps (thread = env) run the policy head, the wrapped env step and the Transition stores,
//                 and publish the next network input through shared memory.
//   PRNG warps    four more warps (thread = env) advance the keys and take the policy's draw for every step while the
//                 producers run the network: the threefry chain never sits on the critical path.
//   issuer        one warp, one lane: 6 x tcgen05.mma (M128 N64 K8, kind::tf32) per slot, tcgen05.commit per layer.
//   few envs      the issue slots of an SM bound a step (the same ~12,000 warp instructions per tile whatever the warp
//                 count), so when the envs do not even fill 64 or 32 rows per SM a CTA takes only that many live rows:
//                 the warps of the other lane quarters leave the step loop (their accumulator rows are never read) and
//                 every barrier is sized to the live warps.
//
// Same arguments, outputs and bits as actor_rollout_tc_kernel.
#pragma once
#include "actor_tc_kernels.cuh"

namespace mbpo {
namespace atcw {

using namespace tc;
using namespace atc;   // TILE, W, QC, QUARTERS, A_LBO_, W_LBO_, SLOT_PLANE, SLOT_BYTES, W_PLANE, MAX_HH, helpers

static_assert(QC == 16 && QUARTERS == 4, "the wide kernel splits a row into four 16-column quarters");
constexpr int PRODUCERS = TILE * QUARTERS;        // 512
constexpr int PRNG_THREADS = TILE;                // 128
constexpr int WTHREADS = PRODUCERS + PRNG_THREADS + 32;   // + the issuer warp = 672
#ifndef MBPO_ATCW_A_TMEM
#define MBPO_ATCW_A_TMEM 0
#endif
// Experiment (-DMBPO_ATCW_A_TMEM=1): A operand in tensor memory -- the producers write the hi / lo planes with
// tcgen05.st next to the accumulators and the MMAs read them from there (no shared-memory stores, no generic -> async
// proxy fence per slice).  Same bits; measured 4.55-4.62 us per step against 4.50-4.58 us through shared memory
// (tcgen05.st + wait::st cost what st.shared + the proxy fence cost), so shared memory stays the default.
constexpr bool A_TMEM = MBPO_ATCW_A_TMEM != 0;
constexpr int WTMEM_COLS = A_TMEM ? 4 * W : 2 * W;   // two 64-column accumulators (+ A hi, A lo: 64 columns each)
constexpr uint32_t TM_A_HI = 2 * W, TM_A_LO = 3 * W;
constexpr int STEP_BAR = 1, PART_BAR = 2, EPS_BAR = 3;    // named barriers

struct WSmem {
  static constexpr uint32_t A = 0;                                         // [slot = quarter][hi, lo]
  static constexpr uint32_t WH = A + QUARTERS * SLOT_BYTES;                // [layer][hi, lo] planes
  static constexpr uint32_t W0 = WH + MAX_HH * 2 * W_PLANE;                // float [3][64]
  static constexpr uint32_t B0 = W0 + 3 * W * 4;                           // float [64]
  static constexpr uint32_t BH = B0 + W * 4;                               // float [MAX_HH][64]
  static constexpr uint32_t WO = BH + MAX_HH * W * 4;                      // float [64][2]
  static constexpr uint32_t BO = WO + W * 2 * 4;                           // float [2] (+ pad)
  static constexpr uint32_t TILES = BO + 16;                               // float [4 warps][96]: row transposition
  static constexpr uint32_t XIN = TILES + 4 * 96 * 4;                      // float4 [128]: the network input of a row
  static constexpr uint32_t PART = XIN + TILE * 16;                        // float2 [3][128]: partials of quarters 1..3
  static constexpr uint32_t EPS = PART + 3 * TILE * 8;                     // float [128]: this step's draw
  static constexpr uint32_t BARS = EPS + TILE * 4;                         // full[4], layer_done
  static constexpr uint32_t TMEM_PTR = BARS + 5 * 8 + 8;
  static constexpr uint32_t TOTAL = TMEM_PTR + 16;
};
static_assert(WSmem::TOTAL <= 227 * 1024, "wide tensor-core actor kernel shared memory plan exceeds 227 KB");

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

#ifdef MBPO_ATCW_PROFILE
#define ATCW_CLK(i)                                  \
  do {                                               \
    const long long now_ = clock64();                \
    prof[i] += now_ - last_;                         \
    last_ = now_;                                    \
  } while (0)
#else
#define ATCW_CLK(i) do { } while (0)
#endif

template <int PRNG, int MATH>
__global__ void __launch_bounds__(WTHREADS, 1) actor_rollout_tc_wide_kernel(const __grid_constant__ ActorArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ float norm_sm[8];     // normaliser mean [0..2], std [4..6]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = tid < PRODUCERS;
  const bool prng_warp = !producer && tid < PRODUCERS + PRNG_THREADS;
  const int q = warp >> 2;                          // column quarter (producers)
  const int r = producer ? (warp & 3) * 32 + lane : (tid - PRODUCERS) & (TILE - 1);   // row of the tile
  float* s_w0 = reinterpret_cast<float*>(smem + WSmem::W0);
  float* s_b0 = reinterpret_cast<float*>(smem + WSmem::B0);
  float* s_bh = reinterpret_cast<float*>(smem + WSmem::BH);
  float* s_wo = reinterpret_cast<float*>(smem + WSmem::WO);
  float* s_bo = reinterpret_cast<float*>(smem + WSmem::BO);
  float4* s_xin = reinterpret_cast<float4*>(smem + WSmem::XIN);
  float2* s_part = reinterpret_cast<float2*>(smem + WSmem::PART);
  float* s_eps = reinterpret_cast<float*>(smem + WSmem::EPS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WSmem::BARS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + WSmem::TMEM_PTR);
  const int L = a.num_hidden, HH = L - 1;

  // ---- one-time setup ----------------------------------------------------------------------------------------
  if (tid < 3) {
    norm_sm[tid] = a.obs_mean_dev ? a.obs_mean_dev[tid] : a.obs_mean[tid];
    norm_sm[4 + tid] = a.obs_std_dev ? a.obs_std_dev[tid] : a.obs_std[tid];
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(WTMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  const int rows = a.rows_per_cta;                  // live rows of this CTA's tile: 32, 64 or 128
  if (tid == 32) {
    for (int i = 0; i < QUARTERS; ++i) mbar_init(&bars[i], QUARTERS * (rows / 32));   // full[j]: every live producer warp
    mbar_init(&bars[QUARTERS], 1);                                        // layer_done
    fence_barrier_init();
  }
  for (int l = 0; l < HH; ++l) {
    const float* wl = a.w[l + 1];                        // flax Dense kernel [in = k][out = n]
    uint8_t* hi = smem + WSmem::WH + (l * 2 + 0) * W_PLANE;
    uint8_t* lo = smem + WSmem::WH + (l * 2 + 1) * W_PLANE;
#pragma unroll 4      // independent loads in flight: the prologue is 15% of a 20-step call
    for (int i = tid; i < W * W; i += WTHREADS) {
      const int k = i / W, n = i % W;
      float h, t;
      split_tf32(wl[i], h, t);
      const uint32_t off = (k / 4) * W_LBO_ + n * 16 + (k % 4) * 4;
      *reinterpret_cast<float*>(hi + off) = h;
      *reinterpret_cast<float*>(lo + off) = t;
    }
    for (int i = tid; i < W; i += WTHREADS) s_bh[l * W + i] = a.b[l + 1][i];
  }
  for (int i = tid; i < 3 * W; i += WTHREADS) s_w0[i] = a.w[0][i];
  for (int i = tid; i < W; i += WTHREADS) s_b0[i] = a.b[0][i];
  for (int i = tid; i < W * 2; i += WTHREADS) s_wo[i] = a.w[L][i];
  if (tid < 2) s_bo[tid] = a.b[L][tid];
  fence_proxy_async();          // the weight planes are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int tile_e0 = blockIdx.x * rows;
  const bool idle = (producer || prng_warp) && r >= rows;    // warp-uniform: a lane quarter without live rows
  const uint32_t bar0 = smem_u32(bars);
  constexpr uint32_t BAR_LAYER_W = QUARTERS * 8u;

  if (!producer && !prng_warp) {
    // ---- MMA issuer ------------------------------------------------------------------------------------------
    if (lane == 0) {
      const uint64_t da0 = umma_desc(smem_u32(smem + WSmem::A), A_LBO_, SBO);
      const uint64_t db0 = umma_desc(smem_u32(smem + WSmem::WH), W_LBO_, SBO);
      constexpr uint32_t IDESC = umma_idesc_tf32(TILE, W);
      uint32_t stage = 0;
#pragma unroll 1
      for (int t = 0; t < a.T; ++t) {
#pragma unroll 1
        for (int s = 0; s < HH; ++s) {
          const uint32_t d = tmem_base + (s & 1) * W;
#pragma unroll 1
          for (int qd = 0; qd < QUARTERS; ++qd) {
            mbar_wait_a(bar0 + qd * 8u, stage & 1u);     // the four warps of quarter qd have written slot qd
            tc_fence_after();
            const uint64_t a_hi = da0 + ((qd * SLOT_BYTES) >> 4), a_lo = a_hi + (SLOT_PLANE >> 4);
            const uint64_t b_hi = db0 + ((static_cast<uint32_t>(s) * 2u * W_PLANE + qd * (QC / 4) * W_LBO_) >> 4);
            const uint64_t b_lo = b_hi + (W_PLANE >> 4);
#pragma unroll
            for (int j = 0; j < QC / 8; ++j) {
              const uint64_t ao = static_cast<uint64_t>((j * 2 * A_LBO_) >> 4), bo = static_cast<uint64_t>((j * 2 * W_LBO_) >> 4);
              if (A_TMEM) {
                const uint32_t col = static_cast<uint32_t>(qd * QC + j * 8);      // K-step: 8 columns of A
                umma_tf32_ts(d, tmem_base + TM_A_LO + col, b_hi + bo, IDESC, (qd | j) ? 1u : 0u);   // small terms first
                umma_tf32_ts(d, tmem_base + TM_A_HI + col, b_lo + bo, IDESC, 1u);
                umma_tf32_ts(d, tmem_base + TM_A_HI + col, b_hi + bo, IDESC, 1u);
              } else {
                umma_tf32_ss(d, a_lo + ao, b_hi + bo, IDESC, (qd | j) ? 1u : 0u);     // small terms first
                umma_tf32_ss(d, a_hi + ao, b_lo + bo, IDESC, 1u);
                umma_tf32_ss(d, a_hi + ao, b_hi + bo, IDESC, 1u);
              }
            }
          }
          umma_commit_a(bar0 + BAR_LAYER_W);             // accumulator s & 1 complete, every slot read
          ++stage;
        }
      }
    }
  } else if (idle) {
    // nothing: the MMAs still cover these rows of the tile, nobody reads them
  } else if (prng_warp) {
    // ---- keys and draws: thread = env row ---------------------------------------------------------------------
    const int e = tile_e0 + r;
    const int ee = e < a.E ? e : a.E - 1;
    Key2 key{a.key_in[0], a.key_in[1]};
    auto advance_keys = [&](Key2 k, Key2& k_act, Key2& k_carry) {
      k_act = k; k_carry = k;
      if (a.key_convention != 2) {
        Key2 first, second;
        split2<PRNG>(k, first, second);
        if (a.key_convention == 0) { k_carry = first; k_act = second; }   // sac.py:290
        else { k_act = first; k_carry = second; }                          // acting.py:70
      }
    };
#pragma unroll 1
    for (int t = 0; t < a.T; ++t) {
      Key2 k_actor, key_next;
      advance_keys(key, k_actor, key_next);
      key = key_next;
      s_eps[r] = actor_draw<PRNG>(a, k_actor, ee);
      named_arrive(EPS_BAR, 2 * rows);                  // the q = 0 warps wait for it before the head
      named_sync(STEP_BAR, 5 * rows);                   // ... and have read it when the step ends
    }
    if (blockIdx.x == 0 && r == 0 && a.key_out) { a.key_out[0] = key.k0; a.key_out[1] = key.k1; }
  } else {
    // ---- producers ----------------------------------------------------------------------------------------------
    const uint32_t tmem_row = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);   // this warp's 32 lanes
    uint32_t layer_phase = 0;
    const PendulumConsts pc(a.sys);
    const float ep_len = static_cast<float>(a.episode_length);
    const float rep = static_cast<float>(a.action_repeat);
    const size_t E = static_cast<size_t>(a.E);
    const int e = tile_e0 + r;
    const bool live = e < a.E;
    const int ee = live ? e : a.E - 1;                 // dead rows shadow the last env, their stores are masked
    const int warp_e0 = e - lane;
    const int rem = a.E - warp_e0;
    const int n_valid = (rem < 32 ? (rem > 0 ? rem : 0) : 32) * 3;
    float* tile = reinterpret_cast<float*>(smem + WSmem::TILES) + (warp & 3) * 96;
    ActorEnv v;
    v.c = a.obs[3 * ee]; v.s = a.obs[3 * ee + 1]; v.w = a.obs[3 * ee + 2];
    v.f_c = a.first_obs[3 * ee]; v.f_s = a.first_obs[3 * ee + 1]; v.f_w = a.first_obs[3 * ee + 2];
    v.steps = a.steps[ee]; v.done = a.done[ee];
    v.th = (MATH == MBPO_MATH_REFERENCE) ? 0.0f : atan2_bounded(v.s, v.c);
    v.f_th = (MATH == MBPO_MATH_REFERENCE) ? 0.0f : atan2_bounded(v.f_s, v.f_c);
    auto network_input = [&](const ActorEnv& s) {
      float x[3] = {s.c, s.s, s.w};
      if (a.normalize) {
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = __fdiv_rn(__fsub_rn(x[i], norm_sm[i]), norm_sm[4 + i]);
      }
      return make_float4(x[0], x[1], x[2], 0.0f);
    };
    float4 xin = network_input(v);      // every quarter starts from the same loaded state
#ifdef MBPO_ATCW_PROFILE
    long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long last_ = clock64();
#endif
    uint8_t* chunk_dst = smem + WSmem::A + q * A_LBO_ + r * 16;   // chunk q of a slot's hi plane (+ slot offset)

#pragma unroll 1
    for (int t = 0; t < a.T; ++t) {
      const f32x2_t x0p = pack2(xin.x, xin.x), x1p = pack2(xin.y, xin.y), x2p = pack2(xin.z, xin.z);
      float pl = 0.0f, ps = 0.0f;       // this quarter's partial of the output layer
#pragma unroll 1
      for (int s = 0; s < HH; ++s) {
        // ---- a stage that feeds MMAs, in four time slices: in slice j the four threads of a row produce the four
        // 4-column chunks of ring slot j (thread q: columns 16 j + 4 q ..), so slot j is complete a quarter of the way
        // through the stage and its six MMAs run under the production of slot j + 1 -- the issuer still takes the
        // slots in the order 0..3, i.e. the accumulation order of the four-tile kernel.
#pragma unroll 1
        for (int j = 0; j < QUARTERS; ++j) {
          const int c0 = j * QC + q * 4;
          float h[4];
          if (s == 0) {
            // layer 0: the same float operations per element as the four-tile kernel
            const float4 r0 = *reinterpret_cast<const float4*>(s_w0 + c0);
            const float4 r1 = *reinterpret_cast<const float4*>(s_w0 + W + c0);
            const float4 r2 = *reinterpret_cast<const float4*>(s_w0 + 2 * W + c0);
            const float4 bb = *reinterpret_cast<const float4*>(s_b0 + c0);
            const float w0r[4] = {r0.x, r0.y, r0.z, r0.w}, w1r[4] = {r1.x, r1.y, r1.z, r1.w};
            const float w2r[4] = {r2.x, r2.y, r2.z, r2.w}, b0r[4] = {bb.x, bb.y, bb.z, bb.w};
            float pre[4];
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
              const f32x2_t acc2 = fma2(x2p, pack2(w2r[i], w2r[i + 1]),
                                        fma2(x1p, pack2(w1r[i], w1r[i + 1]), mul2(x0p, pack2(w0r[i], w0r[i + 1]))));
              unpack2(add2(acc2, pack2(b0r[i], b0r[i + 1])), pre[i], pre[i + 1]);
            }
            swish2(pre[0], pre[1], h[0], h[1]);
            swish2(pre[2], pre[3], h[2], h[3]);
          } else {
            // epilogue of the previous layer: four accumulator columns, bias + swish
            uint32_t acc[4];
            tmem_ld4(tmem_row + ((s - 1) & 1) * W + c0, acc);
            const float4 bb = *reinterpret_cast<const float4*>(s_bh + (s - 1) * W + c0);
            float x0, x1, x2, x3;
            unpack2(add2(pack2(__uint_as_float(acc[0]), __uint_as_float(acc[1])), pack2(bb.x, bb.y)), x0, x1);
            unpack2(add2(pack2(__uint_as_float(acc[2]), __uint_as_float(acc[3])), pack2(bb.z, bb.w)), x2, x3);
            swish2(x0, x1, h[0], h[1]);
            swish2(x2, x3, h[2], h[3]);
          }
          // hi / lo planes of the chunk into slot j (free: layer_done of its last readers was waited for)
          float hi[4], lo[4];
          split_tf32_2(h[0], h[1], hi[0], hi[1], lo[0], lo[1]);
          split_tf32_2(h[2], h[3], hi[2], hi[3], lo[2], lo[3]);
          if (A_TMEM) {
            tmem_st4(tmem_row + TM_A_HI + c0, hi[0], hi[1], hi[2], hi[3]);
            tmem_st4(tmem_row + TM_A_LO + c0, lo[0], lo[1], lo[2], lo[3]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          } else {
            uint8_t* dst = chunk_dst + j * SLOT_BYTES;
            *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(dst + SLOT_PLANE) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            fence_proxy_async();    // generic-proxy writes of A -> visible to the tensor core
          }
          tc_fence_before();        // this thread's tensor-memory accesses are ordered before the MMAs they feed
          __syncwarp();
          if (lane == 0) mbar_arrive_a(bar0 + j * 8u);
        }
        ATCW_CLK(s == 0 ? 0 : 3);                          // production of the stage
        mbar_wait_a(bar0 + BAR_LAYER_W, layer_phase);
        layer_phase ^= 1u;
        tc_fence_after();
        ATCW_CLK(s == 0 ? 2 : 5);                          // MMA wait
      }
      {
        // ---- output layer on the CUDA cores, float32: thread q folds the contiguous columns [16 q, 16 q + 16) of the
        // last hidden layer into a partial, as the four-tile kernel does per quarter
        float h[QC];
        const float* bias = s_bh + (HH - 1) * W + q * QC;
        uint32_t acc[QC];
        tmem_ldq(tmem_row + ((HH - 1) & 1) * W + q * QC, acc);
#pragma unroll
        for (int j4 = 0; j4 < QC / 4; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias + j4 * 4);
          float x0, x1, x2, x3;
          unpack2(add2(pack2(__uint_as_float(acc[j4 * 4]), __uint_as_float(acc[j4 * 4 + 1])), pack2(bb.x, bb.y)), x0, x1);
          unpack2(add2(pack2(__uint_as_float(acc[j4 * 4 + 2]), __uint_as_float(acc[j4 * 4 + 3])), pack2(bb.z, bb.w)), x2, x3);
          swish2(x0, x1, h[j4 * 4], h[j4 * 4 + 1]);
          swish2(x2, x3, h[j4 * 4 + 2], h[j4 * 4 + 3]);
        }
        ATCW_CLK(6);
#pragma unroll
        for (int i = 0; i < QC; i += 2) {
          const float4 wo = *reinterpret_cast<const float4*>(s_wo + (q * QC + i) * 2);
          pl = fmaf(h[i], wo.x, pl);
          ps = fmaf(h[i], wo.y, ps);
          pl = fmaf(h[i + 1], wo.z, pl);
          ps = fmaf(h[i + 1], wo.w, ps);
        }
        ATCW_CLK(7);
      }
      // ---- the four partials of a row meet in its q = 0 thread -----------------------------------------------------
      if (q != 0) {
        s_part[(q - 1) * TILE + r] = make_float2(pl, ps);
        named_arrive(PART_BAR, 4 * rows);
      } else {
        named_sync(PART_BAR, 4 * rows);
        ATCW_CLK(8);
        const float2 p1 = s_part[r], p2 = s_part[TILE + r], p3 = s_part[2 * TILE + r];
        const float loc = __fadd_rn(__fadd_rn(__fadd_rn(pl, p1.x), p2.x), p3.x);
        const float raw_sc = __fadd_rn(__fadd_rn(__fadd_rn(ps, p1.y), p2.y), p3.y);
        named_sync(EPS_BAR, 2 * rows);                  // the PRNG warps have published this step's draw
        ATCW_CLK(9);
        const float eps = s_eps[r];
        // ---- head, wrapped env step, Transition ----------------------------------------------------------------
        const float u = actor_head(a, eps, loc + s_bo[0], raw_sc + s_bo[1], ee, live, t);
        float trunc;
        const float rew = actor_env_step<MATH>(a, pc, v, u, ep_len, rep, trunc);
        s_xin[r] = network_input(v);
        const size_t row = static_cast<size_t>(t) * E;
        if (warp_e0 < a.E)        // warp-uniform
          warp_store3(tile, a.next_observation_out + (row + warp_e0) * 3 + lane, lane, n_valid, v.c, v.s, v.w);
        if (live) {
          a.action_out[row + e] = u;
          a.reward_out[row + e] = rew;
          a.discount_out[row + e] = 1.0f - v.done;
          a.truncation_out[row + e] = trunc;
        }
      }
      ATCW_CLK(10);               // head + env step + stores (q = 0)
      tc_fence_before();          // this step's accumulator reads precede the next step's MMAs
      named_sync(STEP_BAR, 5 * rows);
      xin = s_xin[r];
      ATCW_CLK(11);
    }
#ifdef MBPO_ATCW_PROFILE
    if (blockIdx.x == 0 && (tid == 0 || tid == 3 * TILE))
      printf("atcw q=%d cycles/step: l0 %lld put0 %lld wait0 %lld ep1 %lld put1 %lld wait1 %lld ep_out %lld out %lld "
             "part %lld eps %lld head %lld stepbar %lld\n", q, prof[0] / a.T, prof[1] / a.T, prof[2] / a.T, prof[3] / a.T,
             prof[4] / a.T, prof[5] / a.T, prof[6] / a.T, prof[7] / a.T, prof[8] / a.T, prof[9] / a.T, prof[10] / a.T,
             prof[11] / a.T);
#endif
    if (q == 0 && live) {
      a.obs[3 * e] = v.c; a.obs[3 * e + 1] = v.s; a.obs[3 * e + 2] = v.w;
      a.steps[e] = v.steps;
      a.done[e] = v.done;
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(WTMEM_COLS));
  }
}

}  // namespace atcw
}  // namespace mbpo
