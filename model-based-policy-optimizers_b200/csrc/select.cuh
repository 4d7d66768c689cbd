// Top-K elite selection + mean/std refit + best tracking
// (mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:199-226), executed by one CTA.
//
//   best_elite_idx = argsort(values)[-K:]     stable ascending, IEEE total order   :199
//   elites = action_samples[idx]; elite_values = values[idx]                        :202-203
//   elite_mean / elite_var (ddof=0, two pass)                                       :206-207
//   mean = mean*alpha + (1-alpha)*elite_mean ; var likewise ; std = sqrt(var)       :210-214
//   if best_value <= elite_values[-1]: best = (elite_values[-1], elites[-1])        :217-226
//
// The sort key of element i is the pair (total_order_key(value_i), i); pairs are unique, so
// "top K of a stable ascending sort" is exactly "the K largest pairs".  The K-th largest
// 32-bit key is found by a bitwise binary search in which every thread counts its own
// candidates and the CTA sums the counts (one barrier per differing bit); ties at the
// threshold keep the largest indices.  The K elites are then ranked among themselves (K
// compares each) so that every float sum below runs in the reference's rank order with
// unfused float32 operations: given identical inputs the refit is bit-identical to the
// NumPy oracle.
#pragma once
#include "mathx.cuh"

namespace mbpo {

struct RefitScalars {
  int M;            // N + Np candidates
  int K;            // elites
  int D;            // H * A columns
  float alpha;
  float one_minus_alpha;  // float32(1 - alpha), rounded on the host like the weak python scalar
};

// Scratch words the selection needs besides keys/elite_idx: counters + K selected keys.
__host__ __device__ constexpr int select_scratch_words(int K) { return 40 + K; }

// keys      : shared, uint32[M]  total-order keys of the objective values
// elite_idx : shared, int32[K]   out: argsort(values)[-K:]  (ascending rank)
// sel_idx   : shared, int32[K]   temporary (unordered selection)
// scratch   : shared, uint32[select_scratch_words(K)]  temporary
// row(i, d) : action element d of candidate i
// mean/std/best_seq : shared float[D], updated in place; best_value: shared float*
// Must be called by all THREADS threads of the CTA (contains barriers); on return the
// outputs are visible to every thread.
template <int THREADS, typename RowFn>
__device__ __forceinline__ void cta_select_refit(const RefitScalars rs, const uint32_t* keys, int* elite_idx,
                                                 int* sel_idx, uint32_t* scratch, RowFn row, float* mean,
                                                 float* std_, float* best_seq, float* best_value) {
  const unsigned full = 0xFFFFFFFFu;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int M = rs.M, K = rs.K;
  // scratch layout: [0] or, [1] and, [2] n_gt, [3] n_eq, [4] selection cursor, [8..40) per-bit counts, [40..) keys
  uint32_t* cnt = scratch + 8;
  uint32_t* sel_key = scratch + 40;

  // ---- 0. reset counters; OR / AND of all keys -------------------------------------------
  for (int i = tid; i < 40; i += THREADS) scratch[i] = (i == 1) ? 0xFFFFFFFFu : 0u;
  __syncthreads();
  {
    uint32_t o = 0u, a = 0xFFFFFFFFu;
    for (int i = tid; i < M; i += THREADS) {
      const uint32_t k = keys[i];
      o |= k;
      a &= k;
    }
    o = __reduce_or_sync(full, o);
    a = __reduce_and_sync(full, a);
    if (lane == 0) {
      atomicOr(&scratch[0], o);
      atomicAnd(&scratch[1], a);
    }
  }
  __syncthreads();
  const uint32_t all_and = scratch[1];
  const uint32_t diff = scratch[0] ^ all_and;

  // ---- 1. K-th largest key: bitwise binary search over the bits that differ ----------------
  uint32_t thr = all_and;  // common bits
  for (int bit = 31 - __clz(diff | 1u); bit >= 0; --bit) {
    if (!((diff >> bit) & 1u)) continue;  // CTA-uniform
    const uint32_t cand = thr | (1u << bit);
    int c = 0;
    for (int i = tid; i < M; i += THREADS) c += (keys[i] >= cand) ? 1 : 0;
    c = __reduce_add_sync(full, c);
    if (lane == 0 && c) atomicAdd(&cnt[bit], static_cast<uint32_t>(c));
    __syncthreads();
    if (static_cast<int>(cnt[bit]) >= K) thr = cand;
  }
  // thr is now the K-th largest key (bits below the lowest differing bit are common).
  {
    int n_gt = 0, n_eq = 0;
    for (int i = tid; i < M; i += THREADS) {
      const uint32_t k = keys[i];
      n_gt += (k > thr) ? 1 : 0;
      n_eq += (k == thr) ? 1 : 0;
    }
    n_gt = __reduce_add_sync(full, n_gt);
    n_eq = __reduce_add_sync(full, n_eq);
    if (lane == 0) {
      if (n_gt) atomicAdd(&scratch[2], static_cast<uint32_t>(n_gt));
      if (n_eq) atomicAdd(&scratch[3], static_cast<uint32_t>(n_eq));
    }
  }
  __syncthreads();
  const int skip = static_cast<int>(scratch[3]) - (K - static_cast<int>(scratch[2]));  // ties that do NOT make the cut

  // ---- 2. selection (unordered): k > thr, or a tie whose ascending-index rank is >= skip ---
  for (int i = tid; i < M; i += THREADS) {
    const uint32_t k = keys[i];
    bool sel = k > thr;
    if (k == thr) {
      sel = true;
      if (skip > 0) {  // rare: the cut falls inside a group of equal values
        int tie_rank = 0;
        for (int j = 0; j < i; ++j) tie_rank += (keys[j] == thr) ? 1 : 0;
        sel = tie_rank >= skip;
      }
    }
    if (sel) {
      const uint32_t pos = atomicAdd(&scratch[4], 1u);
      sel_idx[pos] = i;
      sel_key[pos] = k;
    }
  }
  __syncthreads();

  // ---- 3. rank the K elites by (key, index) ascending --------------------------------------
  for (int e = tid; e < K; e += THREADS) {
    const int ie = sel_idx[e];
    const uint32_t ke = sel_key[e];
    int rank = 0;
    for (int f = 0; f < K; ++f) {
      const int jf = sel_idx[f];
      const uint32_t kf = sel_key[f];
      rank += (kf < ke || (kf == ke && jf < ie)) ? 1 : 0;
    }
    elite_idx[rank] = ie;
  }
  __syncthreads();

  // ---- 4. refit, column-parallel, rank-ordered unfused float32 sums ------------------------
  const float kf = static_cast<float>(K);
  const int best_i = elite_idx[K - 1];
  const uint32_t best_key = keys[best_i];
  const float best_elite =
      __uint_as_float((best_key & 0x80000000u) ? (best_key & 0x7FFFFFFFu) : ~best_key);  // invert total_order_key
  const bool take = (*best_value <= best_elite);
  __syncthreads();  // every thread has read *best_value
  for (int d = tid; d < rs.D; d += THREADS) {
    float acc = 0.0f;
#pragma unroll 5
    for (int e = 0; e < K; ++e) acc = __fadd_rn(acc, row(elite_idx[e], d));
    const float emean = __fdiv_rn(acc, kf);
    acc = 0.0f;
#pragma unroll 5
    for (int e = 0; e < K; ++e) {
      const float dl = __fsub_rn(row(elite_idx[e], d), emean);
      acc = __fadd_rn(acc, __fmul_rn(dl, dl));
    }
    const float evar = __fdiv_rn(acc, kf);
    const float m_old = mean[d], s_old = std_[d];
    mean[d] = __fadd_rn(__fmul_rn(m_old, rs.alpha), __fmul_rn(rs.one_minus_alpha, emean));
    const float var = __fadd_rn(__fmul_rn(__fmul_rn(s_old, s_old), rs.alpha), __fmul_rn(rs.one_minus_alpha, evar));
    std_[d] = __fsqrt_rn(var);
    if (take) best_seq[d] = row(best_i, d);
  }
  if (tid == 0 && take) *best_value = best_elite;
  __syncthreads();
}

}  // namespace mbpo
