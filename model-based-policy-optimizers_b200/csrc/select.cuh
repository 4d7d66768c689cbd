// Warp-level top-K elite selection + mean/std refit + best tracking
// (mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:199-226), executed by ONE warp.
//
//   best_elite_idx = argsort(values)[-K:]     stable ascending, IEEE total order   :199
//   elites = action_samples[idx]; elite_values = values[idx]                        :202-203
//   elite_mean / elite_var (ddof=0, two pass)                                       :206-207
//   mean = mean*alpha + (1-alpha)*elite_mean ; var likewise ; std = sqrt(var)       :210-214
//   if best_value <= elite_values[-1]: best = (elite_values[-1], elites[-1])        :217-226
//
// The sort key of element i is the pair (total_order_key(value_i), i); pairs are unique, so
// "top K of a stable ascending sort" is exactly "the K largest pairs".  The K-th largest
// 32-bit key is found by a bitwise binary search with warp ballots; ties at the threshold
// keep the largest indices.  The K elites are then ranked (K^2/32 compares per lane) so that
// every float sum below runs in the reference's rank order with unfused float32 operations:
// given identical inputs the refit is bit-identical to the NumPy oracle.
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

// keys      : shared, uint32[M]  total-order keys of the objective values
// elite_idx : shared, int32[K]   out: argsort(values)[-K:]  (ascending rank)
// scratch   : shared, int32[K]   temporary (index-ordered selection)
// row(i, d) : action element d of candidate i
// mean/std/best_seq : shared float[D], updated in place; best_value: shared float*
template <typename RowFn>
__device__ __forceinline__ void warp_select_refit(const RefitScalars rs, const uint32_t* keys, int* elite_idx,
                                                  int* scratch, RowFn row, float* mean, float* std_, float* best_seq,
                                                  float* best_value) {
  const unsigned full = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int M = rs.M, K = rs.K;
  const int chunks = (M + 31) >> 5;

  // ---- 1. K-th largest key: bitwise binary search over the bits that differ -------------
  uint32_t all_or = 0u, all_and = 0xFFFFFFFFu;
  for (int j = 0; j < chunks; ++j) {
    const int i = lane + (j << 5);
    if (i < M) {
      const uint32_t k = keys[i];
      all_or |= k;
      all_and &= k;
    }
  }
  all_or = __reduce_or_sync(full, all_or);
  all_and = __reduce_and_sync(full, all_and);
  const uint32_t diff = all_or ^ all_and;
  uint32_t thr = all_and;  // common bits
  for (int bit = 31 - __clz(diff | 1u); bit >= 0; --bit) {
    if (!((diff >> bit) & 1u)) continue;
    const uint32_t cand = thr | (1u << bit);
    int cnt = 0;
    for (int j = 0; j < chunks; ++j) {
      const int i = lane + (j << 5);
      cnt += (i < M && keys[i] >= cand) ? 1 : 0;
    }
    cnt = __reduce_add_sync(full, cnt);
    if (cnt >= K) thr = cand;
  }
  // thr is now the K-th largest key (bits below the lowest differing bit are common).
  int n_gt = 0, n_eq = 0;
  for (int j = 0; j < chunks; ++j) {
    const int i = lane + (j << 5);
    if (i < M) {
      const uint32_t k = keys[i];
      n_gt += (k > thr) ? 1 : 0;
      n_eq += (k == thr) ? 1 : 0;
    }
  }
  n_gt = __reduce_add_sync(full, n_gt);
  n_eq = __reduce_add_sync(full, n_eq);
  const int skip = n_eq - (K - n_gt);  // ties (ascending index) that do NOT make the cut

  // ---- 2. compaction in ascending index order --------------------------------------------
  const unsigned lt_mask = (1u << lane) - 1u;
  int base = 0, ties_seen = 0;
  for (int j = 0; j < chunks; ++j) {
    const int i = lane + (j << 5);
    const bool valid = i < M;
    const uint32_t k = valid ? keys[i] : 0u;
    const bool is_eq = valid && (k == thr);
    const unsigned eq_mask = __ballot_sync(full, is_eq);
    const int tie_rank = ties_seen + __popc(eq_mask & lt_mask);
    const bool sel = valid && ((k > thr) || (is_eq && tie_rank >= skip));
    const unsigned sel_mask = __ballot_sync(full, sel);
    if (sel) scratch[base + __popc(sel_mask & lt_mask)] = i;
    base += __popc(sel_mask);
    ties_seen += __popc(eq_mask);
  }
  __syncwarp();

  // ---- 3. rank the K elites by (key, index) ascending -------------------------------------
  for (int e = lane; e < K; e += 32) {
    const int ie = scratch[e];
    const uint32_t ke = keys[ie];
    int rank = 0;
    for (int f = 0; f < K; ++f) {
      const int jf = scratch[f];
      const uint32_t kf = keys[jf];
      rank += (kf < ke || (kf == ke && jf < ie)) ? 1 : 0;
    }
    elite_idx[rank] = ie;
  }
  __syncwarp();

  // ---- 4. refit, column-parallel, rank-ordered unfused float32 sums -----------------------
  const float kf = static_cast<float>(K);
  const int best_i = elite_idx[K - 1];
  const uint32_t best_key = keys[best_i];
  const float best_elite =
      __uint_as_float((best_key & 0x80000000u) ? (best_key & 0x7FFFFFFFu) : ~best_key);  // invert total_order_key
  const bool take = (*best_value <= best_elite);
  __syncwarp();
  for (int d = lane; d < rs.D; d += 32) {
    float acc = 0.0f;
    for (int e = 0; e < K; ++e) acc = __fadd_rn(acc, row(elite_idx[e], d));
    const float emean = __fdiv_rn(acc, kf);
    acc = 0.0f;
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
  if (lane == 0 && take) *best_value = best_elite;
  __syncwarp();
}

}  // namespace mbpo
