// JAX-compatible threefry2x32 PRNG primitives (device side).
//
// Replaces jax.random.{split,random_bits,uniform,normal} as used by the reference at
// mbpo/optimizers/trajectory_optimizers/icem_optimizer.py:123,155,174-180,246 and
// mbpo/utils/general_utils.py:189-191.  Algorithm: Random123 threefry2x32, 20 rounds,
// with JAX's counter layouts (legacy "halves" layout and jax_threefry_partitionable).
#pragma once
#include <stdint.h>

namespace mbpo {

struct Key2 {
  uint32_t k0, k1;
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// One threefry2x32-20 block: (x0, x1) encrypted under (k0, k1).
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define MBPO_TF_ROUND(r) \
  x0 += x1;              \
  x1 = rotl32(x1, r);    \
  x1 ^= x0;
  MBPO_TF_ROUND(13) MBPO_TF_ROUND(15) MBPO_TF_ROUND(26) MBPO_TF_ROUND(6)
  x0 += k1; x1 += k2 + 1u;
  MBPO_TF_ROUND(17) MBPO_TF_ROUND(29) MBPO_TF_ROUND(16) MBPO_TF_ROUND(24)
  x0 += k2; x1 += k0 + 2u;
  MBPO_TF_ROUND(13) MBPO_TF_ROUND(15) MBPO_TF_ROUND(26) MBPO_TF_ROUND(6)
  x0 += k0; x1 += k1 + 3u;
  MBPO_TF_ROUND(17) MBPO_TF_ROUND(29) MBPO_TF_ROUND(16) MBPO_TF_ROUND(24)
  x0 += k1; x1 += k2 + 4u;
  MBPO_TF_ROUND(13) MBPO_TF_ROUND(15) MBPO_TF_ROUND(26) MBPO_TF_ROUND(6)
  x0 += k2; x1 += k0 + 5u;
#undef MBPO_TF_ROUND
}

// Word `w` (0 <= w < n) of the flat legacy output threefry_2x32(key, iota(n)):
// counters are zero padded to even length, cut in halves (x0 = first, x1 = second half),
// and the outputs are concatenated [y0 | y1].  Costs one block.
__device__ __forceinline__ uint32_t legacy_word(Key2 key, uint32_t n, uint32_t w) {
  const uint32_t npad = n + (n & 1u);
  const uint32_t h = npad >> 1;
  const uint32_t j = (w < h) ? w : (w - h);
  uint32_t x0 = j;
  uint32_t x1 = h + j;
  if (x1 >= n) x1 = 0u;  // the padding counter
  threefry2x32(key.k0, key.k1, x0, x1);
  return (w < h) ? x0 : x1;
}

// jax.random.split(key, num)[i]
template <int MODE>
__device__ __forceinline__ Key2 split_at(Key2 key, uint32_t num, uint32_t i) {
  Key2 out;
  if (MODE == 1) {
    uint32_t x0 = 0u, x1 = i;
    threefry2x32(key.k0, key.k1, x0, x1);
    out.k0 = x0;
    out.k1 = x1;
  } else {
    out.k0 = legacy_word(key, 2u * num, 2u * i);
    out.k1 = legacy_word(key, 2u * num, 2u * i + 1u);
  }
  return out;
}

// random_bits(key, (n,))[w]
template <int MODE>
__device__ __forceinline__ uint32_t random_bits_at(Key2 key, uint32_t n, uint32_t w) {
  if (MODE == 1) {
    uint32_t x0 = 0u, x1 = w;
    threefry2x32(key.k0, key.k1, x0, x1);
    return x0 ^ x1;
  }
  return legacy_word(key, n, w);
}

// split(key, 1)[0]: both words of the only child come out of ONE block in either layout (legacy: counters
// (0, 1) -> flat [y0, y1]; partitionable: counter (0, 0) -> (y0, y1)).
template <int MODE>
__device__ __forceinline__ Key2 split1(Key2 key) {
  uint32_t x0 = 0u, x1 = (MODE == 1) ? 0u : 1u;
  threefry2x32(key.k0, key.k1, x0, x1);
  return Key2{x0, x1};
}

// split(key, 2) -> both children.  Legacy: flat = [y0(0,2), y0(1,3), y1(0,2), y1(1,3)].
template <int MODE>
__device__ __forceinline__ void split2(Key2 key, Key2& a, Key2& b) {
  if (MODE == 1) {
    a = split_at<1>(key, 2u, 0u);
    b = split_at<1>(key, 2u, 1u);
  } else {
    uint32_t p0 = 0u, p1 = 2u, q0 = 1u, q1 = 3u;
    threefry2x32(key.k0, key.k1, p0, p1);
    threefry2x32(key.k0, key.k1, q0, q1);
    a.k0 = p0; a.k1 = q0;
    b.k0 = p1; b.k1 = q1;
  }
}

// The two keys powerlaw_psd_gaussian uses out of split(rng, 3) (general_utils.py:189).
// Legacy: flat = [y0_0, y0_1, y0_2, y1_0, y1_1, y1_2] from blocks (0,3), (1,4), (2,5).
template <int MODE>
__device__ __forceinline__ void split3_first2(Key2 key, Key2& a, Key2& b) {
  if (MODE == 1) {
    a = split_at<1>(key, 3u, 0u);
    b = split_at<1>(key, 3u, 1u);
  } else {
    uint32_t p0 = 0u, p1 = 3u, q0 = 1u, q1 = 4u, r0 = 2u, r1 = 5u;
    threefry2x32(key.k0, key.k1, p0, p1);
    threefry2x32(key.k0, key.k1, q0, q1);
    threefry2x32(key.k0, key.k1, r0, r1);
    a.k0 = p0; a.k1 = q0;
    b.k0 = r0; b.k1 = p1;
  }
}

}  // namespace mbpo
