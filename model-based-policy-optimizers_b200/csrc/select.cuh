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
// "top K of a stable ascending sort" is exactly "the K largest pairs".  Selection is one
// radix pass: a 256-bin shared-memory histogram over the 8 most significant bits in which the
// keys differ locates the bin holding the K-th largest key; everything in higher bins is an
// elite, and the (few) candidates inside the boundary bin are ranked exactly by brute force.
// The elites' positions in the ascending order come out of the same comparisons, so every float
// sum below runs in the reference's rank order with unfused float32 operations: given identical
// inputs the refit is bit-identical to the NumPy oracle.
#pragma once
#include "mathx.cuh"

namespace mbpo {

#ifdef MBPO_CLUSTER_CLOCKS
__device__ long long g_select_clocks[16];
#define MBPO_SEL_CLK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_select_clocks[i] = clock64(); } while (0)
#else
#define MBPO_SEL_CLK(i) do {} while (0)
#endif

struct RefitScalars {
  int M;            // N + Np candidates
  int K;            // elites
  int D;            // H * A columns
  float alpha;
  float one_minus_alpha;  // float32(1 - alpha), rounded on the host like the weak python scalar
};

// Scratch words the selection needs besides keys / elite_idx / sel_idx:
// histogram [256 bins, one pad word per 8 bins = 288] + misc [8] + selected keys [K] + boundary-bin candidates [M].
// The pad makes "lane l reads bins 8 l .. 8 l + 7" (step 2) hit 32 different banks instead of 4.
constexpr int SELECT_HIST_WORDS = 288;
constexpr int SELECT_HEAD_WORDS = SELECT_HIST_WORDS + 8;
__host__ __device__ constexpr int select_scratch_words(int K, int M) { return SELECT_HEAD_WORDS + K + M; }
__device__ __forceinline__ uint32_t select_hist_slot(uint32_t bin) { return bin + (bin >> 3); }

// keys      : shared, uint32[M]  total-order keys of the objective values
// elite_idx : shared, int32[K]   out: argsort(values)[-K:]  (ascending rank)
// sel_idx   : shared, int32[K]   temporary (the keys above the boundary bin)
// scratch   : shared, uint32[select_scratch_words(K, M)]  temporary
// row(i, d) : action element d of candidate i
// mean/std/best_seq : shared float[D], updated in place; best_value: shared float*
// Must be called by all THREADS threads of the CTA (contains barriers; THREADS a multiple of
// 32); keys must be visible (barrier) before the call; on return the outputs are visible to
// every thread.
// THREADS == 0: the CTA's size is read from blockDim.x (the cluster plan launches CTAs of N / cluster_size threads).
template <int THREADS>
__device__ __forceinline__ int cta_threads() { return THREADS ? THREADS : static_cast<int>(blockDim.x); }

// Steps 0-5: elite_idx[K] = argsort(values)[-K:] in ascending rank.  On return elite_idx is visible to every thread.
template <int THREADS>
__device__ __forceinline__ void cta_select(const RefitScalars rs, const uint32_t* keys, int* elite_idx, int* sel_idx,
                                           uint32_t* scratch) {
  const int NT = cta_threads<THREADS>();
  const unsigned full = 0xFFFFFFFFu;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int M = rs.M, K = rs.K;
  uint32_t* hist = scratch;                           // [288] padded bins (select_hist_slot)
  uint32_t* misc = scratch + SELECT_HIST_WORDS;       // [0] or, [1] and, [2] selection cursor, [3] candidate cursor
  uint32_t* sel_key = scratch + SELECT_HEAD_WORDS;    // [K]
  int* cand = reinterpret_cast<int*>(scratch + SELECT_HEAD_WORDS + K);  // [M]

  MBPO_SEL_CLK(0);
  // ---- 0. reset; OR / AND of all keys (which bits differ at all) -----------------------------
  for (int i = tid; i < SELECT_HEAD_WORDS; i += NT) scratch[i] = (i == SELECT_HIST_WORDS + 1) ? 0xFFFFFFFFu : 0u;
  __syncthreads();
  {
    uint32_t o = 0u, a = 0xFFFFFFFFu;
    for (int i = tid; i < M; i += NT) {
      const uint32_t k = keys[i];
      o |= k;
      a &= k;
    }
    o = __reduce_or_sync(full, o);
    a = __reduce_and_sync(full, a);
    if (lane == 0) {
      atomicOr(&misc[0], o);
      atomicAnd(&misc[1], a);
    }
  }
  __syncthreads();
  const uint32_t diff = misc[0] ^ misc[1];
  const int hb = 31 - __clz(diff | 1u);        // highest differing bit (0 if all keys are equal)
  const int shift = hb > 7 ? hb - 7 : 0;       // digit = bits [shift, shift + 8); higher bits are common

  MBPO_SEL_CLK(1);
  // ---- 1. histogram of the leading digit ---------------------------------------------------
  for (int i = tid; i < M; i += NT) atomicAdd(&hist[select_hist_slot((keys[i] >> shift) & 255u)], 1u);
  __syncthreads();

  MBPO_SEL_CLK(2);
  // ---- 2. boundary bin: the largest bin bb with count(bins >= bb) >= K (every warp redundantly)
  int bb, n_above;
  {
    uint32_t h8[8];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h8[j] = hist[lane * 9 + j];                    // bins 8 lane .. 8 lane + 7 in the padded layout
      s += h8[j];
    }
    uint32_t suf = s;  // inclusive suffix sum over lanes (lanes >= mine)
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_down_sync(full, suf, off);
      if (lane + off < 32) suf += v;
    }
    const unsigned ok = __ballot_sync(full, static_cast<int>(suf) >= K);  // lane 0 always qualifies (suf = M >= K)
    const int L = 31 - __clz(ok);
    int my_bb = 0, my_above = 0;
    if (lane == L) {
      uint32_t above = suf - s;  // keys in bins above this lane's eight
      my_bb = lane * 8;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (static_cast<int>(above + h8[j]) >= K) {
          my_bb = lane * 8 + j;
          break;
        }
        above += h8[j];
      }
      my_above = static_cast<int>(above);
    }
    bb = __shfl_sync(full, my_bb, L);
    n_above = __shfl_sync(full, my_above, L);
  }

  MBPO_SEL_CLK(3);
  // ---- 3. classify: above the boundary bin -> elite; inside it -> candidate ------------------
  // (a warp reserves its slots of either list with one atomic: the two cursors are the only contended words)
  for (int base = 0; base < M; base += NT) {
    const int i = base + tid;
    uint32_t k = 0u;
    int d = -1;
    if (i < M) {
      k = keys[i];
      d = static_cast<int>((k >> shift) & 255u);
    }
    const unsigned above = __ballot_sync(full, d > bb);
    const unsigned inside = __ballot_sync(full, d == bb);
    uint32_t pos_a = 0u, pos_c = 0u;
    if (lane == 0) {
      if (above) pos_a = atomicAdd(&misc[2], static_cast<uint32_t>(__popc(above)));
      if (inside) pos_c = atomicAdd(&misc[3], static_cast<uint32_t>(__popc(inside)));
    }
    pos_a = __shfl_sync(full, pos_a, 0);
    pos_c = __shfl_sync(full, pos_c, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (d > bb) {
      const uint32_t pos = pos_a + static_cast<uint32_t>(__popc(above & lt));
      sel_idx[pos] = i;
      sel_key[pos] = k;
    } else if (d == bb) {
      cand[pos_c + static_cast<uint32_t>(__popc(inside & lt))] = i;
    }
  }
  __syncthreads();

  MBPO_SEL_CLK(4);
  // ---- 4. ranks.  Every key above the boundary bin outranks every key inside it, so a key above is ranked among the
  // n_above keys above, a key inside among the nc candidates (+ n_above): rank = number of larger (key, index) pairs;
  // rank < K is the elite of ascending position K - 1 - rank.  One pass, four lanes per item (the trip counts are
  // uniform so that every lane reaches the shuffles).
  const int nc = static_cast<int>(misc[3]);
  for (int base = 0; base < 4 * (n_above + nc); base += NT) {
    const int w = base + tid;
    const int item = w >> 2, g = w & 3;
    int larger = 0, ie = 0;
    if (item < n_above) {
      ie = sel_idx[item];
      const uint32_t ke = sel_key[item];
      for (int f = g; f < n_above; f += 4) {
        const int jf = sel_idx[f];
        const uint32_t kf = sel_key[f];
        larger += (kf > ke || (kf == ke && jf > ie)) ? 1 : 0;
      }
    } else if (item < n_above + nc) {
      ie = cand[item - n_above];
      const uint32_t ke = keys[ie];
      larger = (g == 0) ? n_above : 0;
      for (int f = g; f < nc; f += 4) {
        const int jf = cand[f];
        const uint32_t kf = keys[jf];
        larger += (kf > ke || (kf == ke && jf > ie)) ? 1 : 0;
      }
    }
    larger += __shfl_xor_sync(full, larger, 1);
    larger += __shfl_xor_sync(full, larger, 2);
    if (item < n_above + nc && g == 0 && larger < K) elite_idx[K - 1 - larger] = ie;
  }
  __syncthreads();
  MBPO_SEL_CLK(5);
  MBPO_SEL_CLK(6);

}

// The value behind a total-order key (inverse of total_order_key for every non-NaN, non-zero-sign-ambiguous float).
__device__ __forceinline__ float value_of_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Step 6: refit + best tracking from the K elites in ascending rank.  elite(e, d): action element d of the elite of
// rank e; best_elite: the value of the rank K-1 elite (value_of_key(keys[elite_idx[K-1]])).  Column-parallel,
// rank-ordered unfused float32 sums.  Ends with a barrier: the outputs are visible to every thread.
// One column of the refit: mean / std of element d over the K elites in ascending rank, blended with the old values.
template <typename EliteFn>
__device__ __forceinline__ void refit_column(const RefitScalars& rs, EliteFn elite, int d, float m_old, float s_old,
                                             float& m_new, float& s_new) {
  const int K = rs.K;
  const float kf = static_cast<float>(K);
  float acc = 0.0f;
#pragma unroll 5
  for (int e = 0; e < K; ++e) acc = __fadd_rn(acc, elite(e, d));
  const float emean = __fdiv_rn(acc, kf);
  acc = 0.0f;
#pragma unroll 5
  for (int e = 0; e < K; ++e) {
    const float dl = __fsub_rn(elite(e, d), emean);
    acc = __fadd_rn(acc, __fmul_rn(dl, dl));
  }
  const float evar = __fdiv_rn(acc, kf);
  m_new = __fadd_rn(__fmul_rn(m_old, rs.alpha), __fmul_rn(rs.one_minus_alpha, emean));
  const float var = __fadd_rn(__fmul_rn(__fmul_rn(s_old, s_old), rs.alpha), __fmul_rn(rs.one_minus_alpha, evar));
  s_new = __fsqrt_rn(var);
}

template <int THREADS, typename EliteFn>
__device__ __forceinline__ void cta_refit(const RefitScalars rs, float best_elite, EliteFn elite, float* mean,
                                          float* std_, float* best_seq, float* best_value) {
  const int NT = cta_threads<THREADS>();
  const int tid = threadIdx.x;
  const int K = rs.K;
  const bool take = (*best_value <= best_elite);
  __syncthreads();  // every thread has read *best_value
  for (int d = tid; d < rs.D; d += NT) {
    float m_new, s_new;
    refit_column(rs, elite, d, mean[d], std_[d], m_new, s_new);
    mean[d] = m_new;
    std_[d] = s_new;
    if (take) best_seq[d] = elite(K - 1, d);
  }
  if (tid == 0 && take) *best_value = best_elite;
  __syncthreads();
}

// Select + refit of one CTA that holds every candidate row itself (fused plan, staged refit kernel).
template <int THREADS, typename RowFn>
__device__ __forceinline__ void cta_select_refit(const RefitScalars rs, const uint32_t* keys, int* elite_idx,
                                                 int* sel_idx, uint32_t* scratch, RowFn row, float* mean,
                                                 float* std_, float* best_seq, float* best_value) {
  cta_select<THREADS>(rs, keys, elite_idx, sel_idx, scratch);
  const float best_elite = value_of_key(keys[elite_idx[rs.K - 1]]);
  cta_refit<THREADS>(rs, best_elite, [&](int e, int d) { return row(elite_idx[e], d); }, mean, std_, best_seq,
                     best_value);
}

}  // namespace mbpo
