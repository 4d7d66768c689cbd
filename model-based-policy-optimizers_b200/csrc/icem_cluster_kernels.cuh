// Few problems, many SMs: one planning problem spread over a thread-block CLUSTER.
//
// The fused plan (icem_kernels.cuh) gives a problem one CTA = one SM: at B = 1 (the reference's own test,
// tests/test_icemopt.py) 147 of 148 SMs idle and the plan takes as long as one thread needs for its two rows.
// Here a cluster of C CTAs (C = 2, 4, 8, or the non-portable 16) owns the problem; CTA r samples and rolls out
// candidates [r R, (r + 1) R), R = ceil(N / C), and the elite exchange runs through
// distributed shared memory:
//
//   1. every thread pushes the total-order key of its row's objective into ALL C copies of skey[]  (C remote stores)
//   2. cluster.sync()
//   3. every CTA ranks ITS candidates against its full copy of the keys (rank = number of larger (key, index)
//      pairs; rank < K = elite): exactly the stable argsort's top K, without a sort
//   4. the owner of an elite row sends its element d to the CTA that refits column d (d mod C): rank-ordered ebuf[K][H]
//   5. cluster.sync()
//   6. every CTA refits its columns (the reference's rank-ordered sums) and pushes the new mean / std / best_seq of
//      those columns into all C copies
//   7. cluster.sync()
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

constexpr int CLUSTER_MAX = 16;   // 16 is a non-portable size (cudaFuncAttributeNonPortableClusterSizeAllowed)

#ifdef MBPO_CLUSTER_CLOCKS
__device__ long long g_cluster_clocks[64];
#define MBPO_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && it == 1) g_cluster_clocks[i] = clock64(); } while (0)
#else
#define MBPO_CLK(i) do {} while (0)
#endif

// Shared memory of one CTA of the cluster (words).
template <int H>
struct ClusterSmem {
  static constexpr int HS = H | 1;
  static size_t bytes(int R, int N, int Np, int K) {
    const size_t words = static_cast<size_t>(R) * HS          // this CTA's action rows
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
  }
};

// Cooperative sampling of 32 rows by the whole CTA (lane = row, warp = share of the work): used when a CTA owns few
// rows (R <= COOP_MAX_ROWS), where one thread per row would leave the plan waiting on a single thread's ~3,000
// dependent instructions.  Every warp derives the row's keys (6 threefry blocks, redundantly -- the issue slots are
// idle anyway), then the 2 x tasks threefry-block + normal units of the two half spectra are dealt round-robin to
// the warps, and after a barrier each warp evaluates its share of the DFT output groups.  Same device functions and
// per-output operation order as colored_noise_row: same bits.
constexpr int COOP_MAX_ROWS = 96;

// `parts` warps share one chunk of 32 rows (this warp is share `part` of them); warps of different chunks run side
// by side.  Called by every thread of the CTA (barriers inside); `valid` marks the lanes that own a row.
template <int H, int PRNG, typename Emit>
__device__ __forceinline__ void coop_sample_rows(Key2 sampling_rng, int N, int n, bool valid, int part, int parts,
                                                 const float* __restrict__ scale, float* row, Emit emit) {
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
  __syncthreads();   // the rows' staged half spectra are complete
  float sr[S::F], si[S::F];
  if (valid) load_staged<H>(row, sr, si);
  __syncthreads();   // every warp holds them in registers: the outputs may overwrite the rows
  if (valid)
    for (int g = part; g < dft_groups<H>(); g += parts) detail::dft_one_group<H, 0>(g, sr, si, emit);
}

// iCemTO.optimize for ONE problem, executed by the whole cluster.  Same contract as plan_problem(); zero_value is
// the objective of the all-zero kept-elite row (zero_row_value_kernel or, in the closed loop, the caller's own
// rollout).  Trace dumps are written by the owners (values, actions) and by CTA 0 (the rest).
//
// Per iteration: sample + roll out the CTA's rows -> push the keys to every CTA -> cluster.sync -> selection
// (every CTA, redundantly) -> the owner of an elite row sends element d to the CTA that refits column d
// (d mod C) -> cluster.sync -> each CTA refits its columns and pushes mean / std / best_seq of those columns to
// every CTA -> cluster.sync.
template <int H, int PRNG, int MATH>
__device__ __forceinline__ void plan_problem_cluster(const PlanArgs& a, const ClusterCtaSmem<H>& sm,
                                                     const PendulumConsts& pc, const RefitScalars& rs,
                                                     const float* prev_best, Key2 key_in, Key2& key_new, float x_th,
                                                     float x_w, float zero_value, int slot, int slots, int R) {
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
  }
  if (tid == 0) {
    *sm.best_value = __int_as_float(0xFF800000);  // -inf
    Key2 k_opt;
    split2<PRNG>(key_in, k_opt, key_new);
    sm.carry[0] = k_opt.k0;
    sm.carry[1] = k_opt.k1;
  }
  for (int j = N + tid; j < M; j += NT) sm.skey[j] = total_order_key(zero_value);   // kept-elite rows (:192,:245)
  __syncthreads();

  const int n = rank * R + tid;             // the candidate this thread rolls out
  const bool mine = tid < R && n < N;
  float* row = sm.act + static_cast<size_t>(tid < R ? tid : 0) * HS;
  const bool coop = R <= COOP_MAX_ROWS;

  for (int it = 0; it < a.S; ++it) {
    const size_t tslot = static_cast<size_t>(it) * slots + slot;
    // ---- key plumbing (:174-180) ---------------------------------------------------------------------
    MBPO_CLK(0);
    Key2 ck{sm.carry[0], sm.carry[1]}, sampling_rng, particles_rng;
    split2<PRNG>(ck, sampling_rng, particles_rng);
    __syncthreads();  // every thread has read carry
    MBPO_CLK(1);
    if (tid == NT - 1) {
      const Key2 nk = split_at<PRNG>(sampling_rng, static_cast<uint32_t>(N + 1), 0u);   // key = sampling_rng[0]  (:176)
      sm.carry[0] = nk.k0;
      sm.carry[1] = nk.k1;
    }
    // ---- sampling -----------------------------------------------------------------------------------------
    if (coop) {
      // chunks of 32 rows side by side, warps / chunks warps on each
      const int chunks = (R + 31) >> 5, warps = NT >> 5;
      const int parts = warps / chunks;
      const int chunk = (tid >> 5) / parts, part = (tid >> 5) - chunk * parts;
      const int r = (chunk << 5) + (tid & 31);
      const int nn = rank * R + r;
      float* rw = sm.act + static_cast<size_t>(r < R ? r : 0) * HS;
      coop_sample_rows<H, PRNG>(sampling_rng, N, nn, chunk < chunks && r < R && nn < N, part, parts, a.scale, rw,
                                [&](int t, float y) {
        const float v = __fadd_rn(mean[t], __fmul_rn(y, std_[t]));           // :190
        rw[t] = fminf(fmaxf(v, a.u_min), a.u_max);                           // :191
      });
      __syncthreads();   // the rows are complete before their rollout threads read them
    } else if (mine) {
      const Key2 skey_n = split_at<PRNG>(sampling_rng, static_cast<uint32_t>(N + 1), static_cast<uint32_t>(n + 1));
      const Key2 dim_key = split1<PRNG>(skey_n);
      colored_noise_row<H, PRNG>(dim_key, a.scale, row, nullptr, [&](int t, float y) {
        const float v = __fadd_rn(mean[t], __fmul_rn(y, std_[t]));
        row[t] = fminf(fmaxf(v, a.u_min), a.u_max);
      });
    }
    MBPO_CLK(2);
    // ---- rollout of this thread's row; the key goes to every CTA of the cluster -----------------------------
    if (mine) {
      const float ret = rollout_return_th<MATH, true>(pc, x_th, x_w, H, [&](int t) { return row[t]; });
      MBPO_CLK(3);
      const float val = summarize_particles(ret, a.P, a.summarize);
      const uint32_t key = total_order_key(val);
      for (int c = 0; c < C; ++c) cluster.map_shared_rank(sm.skey, c)[n] = key;
      if (a.trace.values) a.trace.values[tslot * M + n] = val;
      if (a.trace.actions) {
        float* dst = a.trace.actions + (tslot * M + n) * H;
        for (int t = 0; t < H; ++t) dst[t] = row[t];
      }
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
    const bool ranked = 2 * rows_here <= NT;
    uint32_t* cnt = sm.sel_scratch;                              // ranked: [rows_here] rank counters
    uint32_t* n_mine = sm.sel_scratch + 299;                     // (the selection's scratch is dead once it returns)
    int* mine_pos = reinterpret_cast<int*>(sm.sel_scratch) + 300; // ... and ascending position e   (<= K entries each)
    if (ranked) {
      for (int i = tid; i < rows_here; i += NT) cnt[i] = 0u;
      if (tid == 0) *n_mine = 0u;
      __syncthreads();
      const int parts = NT / rows_here;                          // >= 2
      const int chunk = (M + parts - 1) / parts;
      for (int w = tid; w < rows_here * parts; w += NT) {
        const int r = w / parts, part = w - r * parts;
        const int i = r < R ? rank * R + r : N + rank + (r - R) * C;
        if (r < R && i >= N) continue;                           // beyond the last real candidate
        const uint32_t ki = sm.skey[i];
        const int j1 = (part + 1) * chunk < M ? (part + 1) * chunk : M;
        uint32_t larger = 0u;
#pragma unroll 8
        for (int j = part * chunk; j < j1; ++j) {
          const uint32_t kj = sm.skey[j];
          larger += (kj > ki || (kj == ki && j > i)) ? 1u : 0u;
        }
        if (larger) atomicAdd(&cnt[r], larger);
      }
      __syncthreads();
      for (int r = tid; r < rows_here; r += NT) {
        const int i = r < R ? rank * R + r : N + rank + (r - R) * C;
        if ((r < R && i >= N) || static_cast<int>(cnt[r]) >= K) continue;
        const int e = K - 1 - static_cast<int>(cnt[r]);
        const uint32_t slot = atomicAdd(n_mine, 1u);
        mine_row[slot] = r;
        mine_pos[slot] = e;
        cluster.map_shared_rank(sm.elite_idx, 0)[e] = i;         // CTA 0 keeps the index list (trace dumps)
        if (e == K - 1)                                          // the best elite's key goes to everybody (:217-226)
          for (int c = 0; c < C; ++c) cluster.map_shared_rank(sm.carry, c)[2] = sm.skey[i];
      }
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
    // ---- elite element (e, d) goes to the CTA that refits column d ---------------------------------------
    {
      const int n_el = static_cast<int>(*n_mine);
      for (int w = tid; w < n_el * H; w += NT) {
        const int q = w / H, d = w - q * H;
        const int r = mine_row[q], e = mine_pos[q];
        const float v = r < R ? sm.act[static_cast<size_t>(r) * HS + d] : 0.0f;   // a kept-elite row: zeros (:192,:245)
        cluster.map_shared_rank(sm.ebuf, d & (C - 1))[e * H + d] = v;          // C is a power of two
      }
    }
    MBPO_CLK(7);
    cluster.sync();   // this CTA's columns of the elite rows are complete; action rows are free again
    MBPO_CLK(8);
    // ---- refit + best tracking of this CTA's columns (:206-226); results to every CTA ----------------------
    const float best_elite = value_of_key(sm.carry[2]);
    const bool take = (*sm.best_value <= best_elite);
    __syncthreads();   // every thread has read *best_value
    if (tid < H && (tid & (C - 1)) == rank) {
      const int d = tid;
      auto elite = [&](int e, int dd) { return sm.ebuf[e * H + dd]; };
      float m_new, s_new;
      refit_column(rs, elite, d, mean[d], std_[d], m_new, s_new);
      const float b_new = elite(K - 1, d);
      for (int c = 0; c < C; ++c) {
        cluster.map_shared_rank(mean, c)[d] = m_new;
        cluster.map_shared_rank(std_, c)[d] = s_new;
        if (take) cluster.map_shared_rank(best_seq, c)[d] = b_new;
      }
    }
    if (tid == 0 && take) *sm.best_value = best_elite;
    cluster.sync();   // mean / std / best_seq are complete everywhere
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
    __syncthreads();
  }
}

// One cluster plans one problem at a time (cluster-stride over problems).  Launched with cluster dimension C along
// x and R = ceil(N / C) rounded up to a warp as the block size; best_value_out holds the zero-row objectives on entry
// (zero_row_value_kernel), like the one-CTA kernel.
template <int H, int PRNG, int MATH>
__global__ void __launch_bounds__(256, 1) icem_plan_cluster_kernel(const __grid_constant__ PlanArgs a, int R) {
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
    const float zero_value = a.best_value_out[b];
    plan_problem_cluster<H, PRNG, MATH>(a, sm, pc, rs, a.best_seq_in + static_cast<size_t>(b) * H, k_in, k_new,
                                        atan2_bounded(x_s, x_c), x_w, zero_value, b, a.B, R);
    cluster.sync();   // every CTA has read the zero-row value; no CTA runs ahead into a lagging CTA's buffers
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
__global__ void __launch_bounds__(256, 1)
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
      // the objective of the all-zero row from this state: the same rollout zero_row_value_kernel runs
      if (tid == 0) {
        const float ret = rollout_return<MATH, true>(pc, x_c, x_s, x_w, H, [](int) { return 0.0f; });
        sm.xs[3] = summarize_particles(ret, a.P, a.summarize);
      }
      __syncthreads();
      const float zero_value = sm.xs[3];
      plan_problem_cluster<H, PRNG, MATH>(a, sm, pc, rs, sm.best_seq, k_in, k_new, atan2_bounded(x_s, x_c), x_w,
                                          zero_value, 0, 1, R);
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
