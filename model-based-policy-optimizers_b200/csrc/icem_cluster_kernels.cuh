// Few problems, many SMs: one planning problem spread over a thread-block CLUSTER.
//
// The fused plan (icem_kernels.cuh) gives a problem one CTA = one SM: at B = 1 (the reference's own test,
// tests/test_icemopt.py) 147 of 148 SMs idle and the plan takes as long as one thread needs for its two rows.
// Here a cluster of C CTAs (C = 2, 4, 8, or the non-portable 16) owns the problem; CTA r owns candidates
// [r R, (r + 1) R), R = ceil(N / C), and the elite exchange runs through distributed shared memory.  Inside a CTA
// the warps are split: the first ceil(R / 32) warps roll out one candidate per thread, the others sample the colored
// noise of the NEXT iteration meanwhile (neither the noise nor the key chain depends on data; only
// clip(mean + noise * std) waits for the refit).  One iteration:
//
//   1. rollout warps: objective of the row -> total-order key into ALL C copies of skey[]  (C remote stores);
//      sampling warps: noise rows of iteration it + 1 into nz[]
//   2. cluster.sync()
//   3. every CTA ranks ITS candidates against its full copy of the keys (rank = number of larger (key, index)
//      pairs; rank < K = elite): exactly the stable argsort's top K, without a sort  (many rows per CTA: every CTA
//      runs the one-CTA selection on its copy instead and picks out its own elites)
//   4. the owner of an elite row sends it to every CTA: rank-ordered ebuf[K][H]
//   5. cluster.sync()
//   6. every CTA refits all columns for itself (the reference's rank-ordered sums on identical inputs: identical
//      mean / std / best everywhere, no third exchange) and applies them to the presampled noise
//
// Every number is produced by the device functions of the one-CTA kernel in the same order (same key tree, same
// noise row, same rollout step, same selection, same rank-ordered refit sums): the results are the one-CTA kernel's
// bits, whatever C (tests/test_gpu_parity.py::test_cluster_plan_bit_identical).
//
// Reference: mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:134-252; tests/test_icemopt.py:19-32.
#pragma once
#include <cooperative_groups.h>

#include "icem_kernels.cuh"

namespace mbpo {

namespace cg = cooperative_groups;

// Ranked selection up to this many candidate rows per warp (measured at 8.25 rows per warp, C = 4: 70 us against 80 us
// with the one-CTA selection; at 16.5, C = 2, the all-pairs count loses).
#ifndef MBPO_RANKED_ROWS_PER_WARP
#define MBPO_RANKED_ROWS_PER_WARP 9
#endif
#ifndef MBPO_QUIET_SCHED
#define MBPO_QUIET_SCHED 1
#endif

constexpr int CLUSTER_MAX = 16;   // 16 is a non-portable size (cudaFuncAttributeNonPortableClusterSizeAllowed)

#ifdef MBPO_CLUSTER_CLOCKS
__device__ long long g_cluster_clocks[64];
#define MBPO_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && it == 1) g_cluster_clocks[i] = clock64(); } while (0)
#else
#define MBPO_CLK(i) do {} while (0)
#endif

// Threads of a CTA: per chunk of 32 rows one rollout warp and up to seven sampling warps, at most 512 threads (128
// registers each: the kernel holds a sampling and a rollout instance and spills below that).
constexpr int COOP_WARPS_PER_CHUNK = 8;
constexpr int CLUSTER_MAX_THREADS = 512;

// Shared memory of one CTA of the cluster (words).
template <int H>
struct ClusterSmem {
  static constexpr int HS = H | 1;
  static size_t bytes(int R, int N, int Np, int K) {
    const size_t words = static_cast<size_t>(R) * HS          // this CTA's action rows
                         + static_cast<size_t>(R) * HS          // next iteration's noise rows
                         + static_cast<size_t>(K) * H          // rank-ordered elite rows (filled by their owners)
                         + (N + Np)                            // sort keys of ALL candidates (filled by their owners)
                         + 3 * H                               // mean, std, best_seq
                         + 2 * K                               // elite_idx, sel_idx
                         + select_scratch_words(K, N + Np)
                         + 12;                                 // best_value, carry key + best key, state key, true state, pad
    return words * 4;
  }
};

template <int H>
struct ClusterCtaSmem {
  float* act;            // [R][HS]
  float* ebuf;           // [K][H]
  uint32_t* skey;        // [M]
  float* mean;
  float* std_;
  float* best_seq;
  int* elite_idx;
  int* sel_idx;
  uint32_t* sel_scratch;
  float* best_value;     // [1]
  uint32_t* carry;       // [3]
  uint32_t* state_key;   // [2]
  float* xs;             // [4] true state of the closed loop
  float* nz;             // [R][HS] colored noise of the NEXT iteration
  __device__ __forceinline__ ClusterCtaSmem(uint32_t* base, int R, int N, int Np, int K) {
    constexpr int HS = ClusterSmem<H>::HS;
    act = reinterpret_cast<float*>(base);
    ebuf = act + static_cast<size_t>(R) * HS;
    skey = reinterpret_cast<uint32_t*>(ebuf + static_cast<size_t>(K) * H);
    mean = reinterpret_cast<float*>(skey + (N + Np));
    std_ = mean + H;
    best_seq = std_ + H;
    elite_idx = reinterpret_cast<int*>(best_seq + H);
    sel_idx = elite_idx + K;
    sel_scratch = reinterpret_cast<uint32_t*>(sel_idx + K);
    best_value = reinterpret_cast<float*>(sel_scratch + select_scratch_words(K, N + Np));
    carry = reinterpret_cast<uint32_t*>(best_value + 1);   // [0..1] carry key, [2] key of the best elite of the iteration
    state_key = carry + 3;
    xs = reinterpret_cast<float*>(state_key + 2);
    nz = best_value + 12;
  }
};

// Cooperative sampling of 32 rows by several warps (lane = row, warp = share of the work): one thread per row would
// leave the plan waiting on a single thread's ~3,000 dependent instructions.  Every warp derives the row's keys (6 threefry blocks, redundantly -- the issue slots are
// idle anyway), then the 2 x tasks threefry-block + normal units of the two half spectra are dealt round-robin to
// the warps, and after a barrier each warp evaluates its share of the DFT output groups.  Same device functions and
// per-output operation order as colored_noise_row: same bits.
//
// `parts` warps share one chunk of 32 rows (this warp is share `part` of them); warps of different chunks run side
// by side.  Called by every thread of the CTA (barriers inside); `valid` marks the lanes that own a row.
// `bar` is a barrier over exactly the threads that make the call (all of the CTA, or the sampling warps alone).
template <int H, int PRNG, typename Emit, typename Bar>
__device__ __forceinline__ void coop_sample_rows(Key2 sampling_rng, int N, int n, bool valid, int part, int parts,
                                                 const float* __restrict__ scale, float* row, Emit emit, Bar bar) {
  using S = NoiseShape<H>;
  constexpr int TASKS = noise_tasks<H, PRNG>();
  if (valid) {
    const Key2 skey_n = split_at<PRNG>(sampling_rng, static_cast<uint32_t>(N + 1), static_cast<uint32_t>(n + 1));
    const Key2 dim_key = split1<PRNG>(skey_n);
    Key2 key_sr, key_si;
    split3_first2<PRNG>(dim_key, key_sr, key_si);
    for (int task = part; task < 2 * TASKS; task += parts) {
      if (task < TASKS) stage_normals_task<H, PRNG, false>(key_sr, scale, row, nullptr, task);
      else stage_normals_task<H, PRNG, true>(key_si, scale, row, nullptr, task - TASKS);
    }
  }
  bar();   // the rows' staged half spectra are complete
  float sr[S::F], si[S::F];
  if (valid) load_staged<H>(row, sr, si);
  bar();   // every warp holds them in registers: the outputs may overwrite the rows
  if (valid)
    for (int g = part; g < dft_groups<H>(); g += parts) detail::dft_one_group<H, 0>(g, sr, si, emit);
}

// iCemTO.optimize for ONE problem, executed by the whole cluster.  Same contract as plan_problem(); the objective
// of the all-zero kept-elite rows (:192,:245) is rolled out here by one thread, beside the sampling of iteration 0.
// Trace dumps are written by the owners (values, actions) and by CTA 0 (the rest).
template <int H, int PRNG, int MATH>
__device__ __forceinline__ void plan_problem_cluster(const PlanArgs& a, const ClusterCtaSmem<H>& sm,
                                                     const PendulumConsts& pc, const RefitScalars& rs,
                                                     const float* prev_best, Key2 key_in, Key2& key_new, float x_th,
                                                     float x_w, int slot, int slots, int R) {
  constexpr int HS = ClusterSmem<H>::HS;
  cg::cluster_group cluster = cg::this_cluster();
  const int C = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int N = a.N, M = a.N + a.Np, K = a.K;
  const int tid = threadIdx.x, NT = blockDim.x;
  float* mean = sm.mean;
  float* std_ = sm.std_;
  float* best_seq = sm.best_seq;

  // ---- prologue (icem_optimizer.py:235-249), redundantly in every CTA --------------------------------
  float m0 = 0.0f;
  if (tid < H && a.warm_start) m0 = prev_best[tid + 1 < H ? tid + 1 : H - 1];
  __syncthreads();  // prev_best may alias best_seq
  if (tid < H) {
    mean[tid] = m0;
    std_[tid] = a.init_std;
    best_seq[tid] = m0;
    sm.act[tid] = 0.0f;                             // row 0 = the all-zero kept-elite row, rolled out by thread 0 below
  }
  if (tid == 0) {
    *sm.best_value = __int_as_float(0xFF800000);  // -inf
    sm.sel_scratch[299] = 0u;                       // the elite list cursor of the selection
  }
  // The carry key (:174-180) does not depend on data: every thread advances its own copy in registers.
  Key2 chain;
  split2<PRNG>(key_in, chain, key_new);
  auto next_sampling_rng = [&]() {
    Key2 srng, particles_rng;
    split2<PRNG>(chain, srng, particles_rng);
    chain = split_at<PRNG>(srng, static_cast<uint32_t>(N + 1), 0u);   // key = sampling_rng[0]  (:176)
    return srng;
  };
  __syncthreads();

  // ---- who does what.  The colored noise does not depend on mean / std either, so the CTA is split: the first
  // `chunks` warps own one candidate per thread and roll it out; the other warps sample the noise of iteration
  // it + 1 meanwhile (into nz); only  clip(mean + noise * std)  (:190-191) waits for the refit.  `sparts` sampling
  // warps share each chunk of 32 rows (coop_sample_rows; with many rows per CTA sparts is 1: a sampling thread per
  // row).  Pass it = -1 samples iteration 0 and rolls out the all-zero row.
  const int warp = tid >> 5, lane = tid & 31;
  const int chunks = (R + 31) >> 5;                        // rollout warps (32 rows each)
  const int warps = NT >> 5;
  // One or two rollout warps (R <= 64): the warps that share a scheduler with them (warp & 3 < chunks) stay idle in
  // the loop -- all of them with one rollout warp, all but one per scheduler with two (warps 4 and 5) -- so that the
  // rollout chains, the critical path, issue (almost) alone while enough warps remain to hide the sampling under them.
  const bool quiet = MBPO_QUIET_SCHED && chunks <= 2 && warps == 8 * chunks;
  auto samples = [&](int w) {
    if (w < chunks) return false;
    if (!quiet || (w & 3) >= chunks) return true;
    return chunks == 2 && (w >> 2) == 1;
  };
  int n_samplers = 0, sw_mine = 0;
  for (int w = chunks; w < warps; ++w) {
    if (samples(w)) {
      if (w < warp) ++sw_mine;
      ++n_samplers;
    }
  }
  const int sparts = n_samplers / chunks;                  // sampling warps per chunk of 32 rows (host: >= 1)
  const int swarps = chunks * sparts;
  const int n = rank * R + tid;             // the candidate this thread rolls out
  const bool mine = tid < R && n < N;
  float* row = sm.act + static_cast<size_t>(tid < R ? tid : 0) * HS;
  float zero_value = 0.0f;

  for (int it = -1; it < a.S; ++it) {
    const size_t tslot = static_cast<size_t>(it < 0 ? 0 : it) * slots + slot;
    MBPO_CLK(0);
    MBPO_CLK(1);
    if (warp >= chunks) {
      // ---- sampling warps: the noise of iteration it + 1 -----------------------------------------------------
      if (it + 1 < a.S) {
        const int sw = samples(warp) ? sw_mine : swarps;
        if (sw < swarps) {
          const Key2 srng = next_sampling_rng();
          const int chunk = sw / sparts;
          const int r = (chunk << 5) + lane;
          const int nn = rank * R + r;
          float* nrow = sm.nz + static_cast<size_t>(r < R ? r : 0) * HS;
          if (sparts > 1) {
            const int count = swarps << 5;
            coop_sample_rows<H, PRNG>(srng, N, nn, r < R && nn < N, sw - chunk * sparts, sparts, a.scale, nrow,
                                      [&](int t, float y) { nrow[t] = y; },
                                      [count] { asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory"); });
          } else if (r < R && nn < N) {
            // a sampling thread per row: the unrolled single-thread routine beats the shared one run by one share
            const Key2 skey_n = split_at<PRNG>(srng, static_cast<uint32_t>(N + 1), static_cast<uint32_t>(nn + 1));
            const Key2 dim_key = split1<PRNG>(skey_n);
            colored_noise_row<H, PRNG>(dim_key, a.scale, nrow, nullptr, [&](int t, float y) { nrow[t] = y; });
          }
        }
      }
    } else if (it < 0 ? tid == 0 : mine) {
      // ---- rollout of this thread's row; the key goes to every CTA of the cluster -----------------------------
      MBPO_CLK(2);
      const float ret = rollout_return_th<MATH, true>(pc, x_th, x_w, H, [&](int t) { return row[t]; });
      MBPO_CLK(3);
      const float val = summarize_particles(ret, a.P, a.summarize);
      if (it < 0) {
        sm.xs[3] = val;                     // the objective of the kept-elite rows (:192,:245)
      } else {
        const uint32_t key = total_order_key(val);
        for (int c = 0; c < C; ++c) cluster.map_shared_rank(sm.skey, c)[n] = key;
        if (a.trace.values) a.trace.values[tslot * M + n] = val;
        if (a.trace.actions) {
          float* dst = a.trace.actions + (tslot * M + n) * H;
          for (int t = 0; t < H; ++t) dst[t] = row[t];
        }
      }
    }
    if (it < 0) {
      __syncthreads();
      zero_value = sm.xs[3];
      for (int j = N + tid; j < M; j += NT) sm.skey[j] = total_order_key(zero_value);
      // :190-191 on the presampled noise (ends with a barrier)
      for (int w = tid; w < R * H; w += NT) {
        const int r = w / H, t = w - r * H;
        const float v = __fadd_rn(mean[t], __fmul_rn(sm.nz[static_cast<size_t>(r) * HS + t], std_[t]));
        sm.act[static_cast<size_t>(r) * HS + t] = fminf(fmaxf(v, a.u_min), a.u_max);
      }
      __syncthreads();
      continue;
    }
    MBPO_CLK(4);
    cluster.sync();   // all keys are in every copy
    MBPO_CLK(5);
    if (rank == 0 && (a.trace.values || a.trace.actions)) {
      for (int j = N + tid; j < M; j += NT) {
        if (a.trace.values) a.trace.values[tslot * M + j] = zero_value;
        if (a.trace.actions) {
          float* dst = a.trace.actions + (tslot * M + j) * H;
          for (int t = 0; t < H; ++t) dst[t] = 0.0f;
        }
      }
    }
    // ---- selection, distributed (:199-203): every CTA ranks ITS candidates against all keys ------------------
    // rank(i) = number of candidates whose (key, index) pair is larger; the K elites are rank < K and the elite of
    // ascending position e has rank K - 1 - e.  The pairs are distinct, so this is exactly argsort(values)[-K:]
    // (stable, ascending) -- without the histogram passes of cta_select, whose barrier phases cost 6,300 cycles.
    // The Np kept-elite rows N .. M-1 are dealt round-robin to the CTAs as virtual candidates (all-zero actions).
    int* mine_row = sm.sel_idx;                                  // compact list of this CTA's elites: local row ...
    const int V = (a.Np + C - 1 - rank) / C;                     // virtual rows of this CTA: j = N + rank + v * C
    const int rows_here = R + V;                                 // local candidate slots [0, R) real, [R, R + V) virtual
    // With many rows per CTA the all-pairs count costs more than cta_select's passes: then every CTA selects on its
    // full copy of the keys (redundantly, identical results) and picks its own elites out of the list.
    // Ranked: a warp takes RB rows at a time (independent compare chains); its lanes stride over the keys and one
    // __reduce_add per row counts the larger pairs -- no shared counters, no barrier inside.  The loops stay rolled:
    // this code runs once per iteration on a few warps, so every instruction is an instruction-cache miss and a
    // compact loop beats its unrolled form (measured: 3,400 cycles unrolled over 20 keys per lane).
    // (the choice must be the same in every CTA of the cluster -- the ranked CTAs send the best elite's key to all,
    // the others read it from their own list -- so it is made on the largest row count, rank 0's)
    const bool ranked = (R + (a.Np + C - 1) / C) <= MBPO_RANKED_ROWS_PER_WARP * warps;
    uint32_t* n_mine = sm.sel_scratch + 299;                     // (the selection's scratch is dead once it returns)
    int* mine_pos = reinterpret_cast<int*>(sm.sel_scratch) + 300; // ... and ascending position e   (<= K entries each)
    if (ranked) {
      MBPO_CLK(10);
      constexpr int RB = 6;
#pragma unroll 1
      for (int r0 = warp; r0 < rows_here; r0 += RB * warps) {
        unsigned long long pi[RB];
        uint32_t larger[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          const int r = r0 + b * warps;
          const int i = r < R ? rank * R + r : N + rank + (r - R) * C;
          const bool ok = r < rows_here && !(r < R && i >= N);   // warp-uniform
          // (key, index) pairs compared as one 64-bit integer: branch-free, two compares per pair
          pi[b] = ok ? (static_cast<unsigned long long>(sm.skey[i]) << 32) | static_cast<uint32_t>(i) : ~0ull;
          larger[b] = 0u;
        }
#pragma unroll 4
        for (int j = lane; j < M; j += 32) {
          const unsigned long long pj = (static_cast<unsigned long long>(sm.skey[j]) << 32) | static_cast<uint32_t>(j);
#pragma unroll
          for (int b = 0; b < RB; ++b) larger[b] += pj > pi[b] ? 1u : 0u;
        }
        MBPO_CLK(11);
        // lane b finishes row b of the batch
        unsigned long long my_pi = ~0ull;
        uint32_t my_cnt = 0u;
#pragma unroll
        for (int b = 0; b < RB; ++b) {
          const uint32_t cnt = __reduce_add_sync(0xFFFFFFFFu, larger[b]);
          if (lane == b) { my_cnt = cnt; my_pi = pi[b]; }
        }
        if (my_pi != ~0ull && static_cast<int>(my_cnt) < K) {    // (a real pair never has index 2^32 - 1)
          const int r = r0 + lane * warps;
          const uint32_t ki = static_cast<uint32_t>(my_pi >> 32);
          const int i = static_cast<int>(static_cast<uint32_t>(my_pi));
          const int e = K - 1 - static_cast<int>(my_cnt);
          const uint32_t slot = atomicAdd(n_mine, 1u);
          mine_row[slot] = r;
          mine_pos[slot] = e;
          cluster.map_shared_rank(sm.elite_idx, 0)[e] = i;       // CTA 0 keeps the index list (trace dumps)
          if (e == K - 1)                                        // the best elite's key goes to everybody (:217-226)
            for (int c = 0; c < C; ++c) cluster.map_shared_rank(sm.carry, c)[2] = ki;
        }
      }
      MBPO_CLK(12);
    } else {
      cta_select<0>(rs, sm.skey, sm.elite_idx, sm.sel_idx, sm.sel_scratch);   // sel_idx is free again afterwards
      if (tid == 0) {
        *n_mine = 0u;
        sm.carry[2] = sm.skey[sm.elite_idx[K - 1]];
      }
      __syncthreads();
      const uint32_t inv_r = ((1u << 20) + R - 1) / R;           // src / R == (src * inv_r) >> 20 (src < 2^11, R <= 2^8)
      for (int e = tid; e < K; e += NT) {
        const int src = sm.elite_idx[e];
        int r = -1;
        if (src >= N) {
          if (((src - N) & (C - 1)) == rank) r = R + (src - N - rank) / C;
        } else if (static_cast<int>((static_cast<uint32_t>(src) * inv_r) >> 20) == rank) {
          r = src - rank * R;
        }
        if (r >= 0) {
          const uint32_t slot = atomicAdd(n_mine, 1u);
          mine_row[slot] = r;
          mine_pos[slot] = e;
        }
      }
    }
    __syncthreads();
    MBPO_CLK(6);
    // ---- every elite row goes to every CTA (rank-ordered ebuf) ------------------------------------------------
    {
      const int n_el = static_cast<int>(*n_mine);
      for (int w = tid; w < n_el * H; w += NT) {
        const int q = w / H, d = w - q * H;
        const int r = mine_row[q], e = mine_pos[q];
        const float v = r < R ? sm.act[static_cast<size_t>(r) * HS + d] : 0.0f;   // a kept-elite row: zeros (:192,:245)
        for (int c = 0; c < C; ++c) cluster.map_shared_rank(sm.ebuf, c)[e * H + d] = v;
      }
    }
    MBPO_CLK(7);
    cluster.sync();   // the elite rows are complete in every CTA; action rows are free again
    MBPO_CLK(8);
    // ---- refit + best tracking (:206-226), every CTA for itself: identical inputs, identical results, and no
    // third exchange.  (A CTA that runs ahead cannot disturb a slower one: it writes skey / ebuf / carry[2] of the
    // next iteration only after that iteration's first cluster.sync, which the slower CTA reaches after this refit.)
    cta_refit<0>(rs, value_of_key(sm.carry[2]), [&](int e, int d) { return sm.ebuf[e * H + d]; }, mean, std_,
                 best_seq, sm.best_value);
    MBPO_CLK(9);
    if (rank == 0) {
      if (a.trace.elite_idx)
        for (int e = tid; e < K; e += NT) a.trace.elite_idx[tslot * K + e] = sm.elite_idx[e];
      for (int d = tid; d < H; d += NT) {
        if (a.trace.mean) a.trace.mean[tslot * H + d] = mean[d];
        if (a.trace.std) a.trace.std[tslot * H + d] = std_[d];
      }
      if (a.trace.best_value && tid == 0) a.trace.best_value[tslot] = *sm.best_value;
    }
    if (tid == 0) *n_mine = 0u;
    if (it + 1 < a.S) {   // the next iteration's rows (their noise was sampled under the rollouts), :190-191
      for (int w = tid; w < R * H; w += NT) {
        const int r = w / H, t = w - r * H;
        const float v = __fadd_rn(mean[t], __fmul_rn(sm.nz[static_cast<size_t>(r) * HS + t], std_[t]));
        sm.act[static_cast<size_t>(r) * HS + t] = fminf(fmaxf(v, a.u_min), a.u_max);
      }
    }
    __syncthreads();
  }
}

// One cluster plans one problem at a time (cluster-stride over problems).  Launched with cluster dimension C along
// x (the all-zero row's objective is rolled out inside; best_value_out is output only).
template <int H, int PRNG, int MATH>
__global__ void __launch_bounds__(CLUSTER_MAX_THREADS, 1) icem_plan_cluster_kernel(const __grid_constant__ PlanArgs a, int R) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const ClusterCtaSmem<H> sm(smem_u32, R, a.N, a.Np, a.K);
  const int tid = threadIdx.x;
  const PendulumConsts pc(a.sys);
  RefitScalars rs;
  rs.M = a.N + a.Np; rs.K = a.K; rs.D = H; rs.alpha = a.alpha; rs.one_minus_alpha = a.one_minus_alpha;
  const int clusters = static_cast<int>(gridDim.x) / C;
  for (int b = static_cast<int>(blockIdx.x) / C; b < a.B; b += clusters) {
    const float x_c = a.x0[3 * b], x_s = a.x0[3 * b + 1], x_w = a.x0[3 * b + 2];
    Key2 k_in{a.key_in[2 * b], a.key_in[2 * b + 1]}, k_new;
    plan_problem_cluster<H, PRNG, MATH>(a, sm, pc, rs, a.best_seq_in + static_cast<size_t>(b) * H, k_in, k_new,
                                        atan2_bounded(x_s, x_c), x_w, b, a.B, R);
    cluster.sync();   // no CTA runs ahead into the next problem while a lagging CTA still reads this one's buffers
    if (rank == 0) {
      if (tid < H) a.best_seq_out[static_cast<size_t>(b) * H + tid] = sm.best_seq[tid];
      if (tid == 0) {
        a.best_value_out[b] = *sm.best_value;
        a.key_out[2 * b] = k_new.k0;
        a.key_out[2 * b + 1] = k_new.k1;
      }
    }
    __syncthreads();
  }
}

// Closed-loop MPC (tests/test_icemopt.py:19-32) on a cluster: T x { plan; true System.step; warm start }.  Every CTA
// keeps its own copy of the true state, the planner key and the best sequence (identical by construction); CTA 0
// writes the outputs.  The zero-row objective of every step is one rollout by one thread per CTA.
template <int H, int PRNG, int MATH>
__global__ void __launch_bounds__(CLUSTER_MAX_THREADS, 1)
    icem_mpc_cluster_kernel(const __grid_constant__ PlanArgs a, const __grid_constant__ MpcArgs m, int R) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const ClusterCtaSmem<H> sm(smem_u32, R, a.N, a.Np, a.K);
  const int tid = threadIdx.x;
  const PendulumConsts pc(a.sys);
  RefitScalars rs;
  rs.M = a.N + a.Np; rs.K = a.K; rs.D = H; rs.alpha = a.alpha; rs.one_minus_alpha = a.one_minus_alpha;
  const int clusters = static_cast<int>(gridDim.x) / C;
  for (int b = static_cast<int>(blockIdx.x) / C; b < a.B; b += clusters) {
    if (tid < H) sm.best_seq[tid] = a.best_seq_in[static_cast<size_t>(b) * H + tid];
    if (tid < 3) sm.xs[tid] = a.x0[3 * b + tid];
    if (tid == 0) { sm.state_key[0] = a.key_in[2 * b]; sm.state_key[1] = a.key_in[2 * b + 1]; }
    __syncthreads();
    for (int t = 0; t < m.T; ++t) {
      const float x_c = sm.xs[0], x_s = sm.xs[1], x_w = sm.xs[2];
      Key2 k_in{sm.state_key[0], sm.state_key[1]}, k_new;
      plan_problem_cluster<H, PRNG, MATH>(a, sm, pc, rs, sm.best_seq, k_in, k_new, atan2_bounded(x_s, x_c), x_w,
                                          0, 1, R);
      if (tid == 0) {
        sm.state_key[0] = k_new.k0; sm.state_key[1] = k_new.k1;
        const float u = sm.best_seq[0];                       // opt_state.action (:67-69)
        float c = x_c, s = x_s, w = x_w, r;
        if (MATH == MBPO_MATH_REFERENCE) {
          pendulum_step_ref(pc, c, s, w, u, r);
        } else {
          float th = atan2_bounded(s, c);
          pendulum_step_theta(pc, th, w, u, r);
          sincos_bounded(th, s, c);
        }
        sm.xs[0] = c; sm.xs[1] = s; sm.xs[2] = w;
        if (rank == 0) {
          const size_t o = static_cast<size_t>(t) * a.B + b;
          if (m.states_out) { m.states_out[o * 3] = c; m.states_out[o * 3 + 1] = s; m.states_out[o * 3 + 2] = w; }
          if (m.rewards_out) m.rewards_out[o] = r;
          if (m.actions_out) m.actions_out[o] = u;
        }
      }
      __syncthreads();
    }
    cluster.sync();
    if (rank == 0) {
      if (tid < H) a.best_seq_out[static_cast<size_t>(b) * H + tid] = sm.best_seq[tid];
      if (tid == 0) {
        a.best_value_out ? (void)(a.best_value_out[b] = *sm.best_value) : (void)0;
        a.key_out[2 * b] = sm.state_key[0];
        a.key_out[2 * b + 1] = sm.state_key[1];
      }
    }
    __syncthreads();
  }
}

}  // namespace mbpo
