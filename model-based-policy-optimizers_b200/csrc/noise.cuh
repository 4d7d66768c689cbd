// Colored (power-law) Gaussian noise row: powerlaw_psd_gaussian(exponent, H, rng)
// (mbpo/utils/general_utils.py:81-208), one row per thread.
//
//   key_sr, key_si, _ = split(rng, 3)                         :189
//   sr = normal(key_sr, (F,)) * s_scale ; si likewise         :190-191
//   Nyquist (even H) and DC: imaginary part 0, real part * sqrt(2)   :195-201
//   y = irfft(sr + i*si, n=H) / sigma                         :207
//
// The irfft is a direct real DFT with compile-time twiddles (they fold into FFMA immediates)
// that exploits y[t], y[H-t] sharing the cosine sum and having opposite sine sums.
#pragma once
#include "mathx.cuh"
#include "threefry.cuh"

namespace mbpo {

namespace detail {
constexpr double csin(double x) {
  double t = x, s = x;
  for (int n = 1; n < 30; ++n) {
    t *= -x * x / ((2 * n) * (2 * n + 1));
    s += t;
  }
  return s;
}
constexpr double ccos(double x) {
  double t = 1, s = 1;
  for (int n = 1; n < 30; ++n) {
    t *= -x * x / ((2 * n - 1) * (2 * n));
    s += t;
  }
  return s;
}
// cos/sin(2*pi*j/H) * weight/H folded: c[j], s[j] for j in [0, H)
template <int H>
struct Twiddle {
  float c[H];
  float s[H];
  constexpr Twiddle() : c(), s() {
    for (int j = 0; j < H; ++j) {
      double a = 6.283185307179586476925 * j / H;
      if (a > 3.14159265358979323846) a -= 6.283185307179586476925;
      c[j] = static_cast<float>(ccos(a) / H);
      s[j] = static_cast<float>(csin(a) / H);
    }
  }
};
}  // namespace detail

template <int H>
struct NoiseShape {
  static constexpr int F = H / 2 + 1;          // rfft bins
  static constexpr int FPAD = F + (F & 1);     // legacy counter padding
  static constexpr int HALF = FPAD / 2;        // threefry blocks per normal vector (legacy)
  static constexpr bool EVEN = (H % 2) == 0;
};

// Stages the scaled normals of one spectrum half (real or imaginary part) into `stage`:
//   stage[k]         = sr[k]            k in [0, F)
//   stage[F + k - 1] = si[k]            k in [1, LASTK)   (DC and, for even H, Nyquist have no
//                                                           imaginary part: general_utils.py:195-201)
// so the H needed values fill exactly the H floats of the caller's row.  The threefry blocks run
// in a rolled loop (small code: the instruction cache matters more than the last bit of ILP);
// `scale` is indexed dynamically (constant bank / shared memory).
template <int H, int MODE, bool IMAG>
__device__ __forceinline__ void stage_normals(Key2 key, const float* __restrict__ scale, float* stage,
                                              uint32_t* bits_out) {
  using S = NoiseShape<H>;
  constexpr int LASTK = S::EVEN ? S::F - 1 : S::F;
  auto put = [&](int k, uint32_t word) {
    if (bits_out) bits_out[k] = word;
    if (!IMAG) {
      stage[k] = bits_to_normal(word) * scale[k];
    } else if (k >= 1 && k < LASTK) {
      stage[S::F + k - 1] = bits_to_normal(word) * scale[k];
    }
  };
  if (MODE == 1) {
#pragma unroll 2
    for (int f = 0; f < S::F; ++f) {
      uint32_t x0 = 0u, x1 = static_cast<uint32_t>(f);
      threefry2x32(key.k0, key.k1, x0, x1);
      put(f, x0 ^ x1);
    }
  } else {
#pragma unroll 2
    for (int j = 0; j < S::HALF; ++j) {
      uint32_t x0 = static_cast<uint32_t>(j);
      uint32_t x1 = (S::HALF + j < S::F) ? static_cast<uint32_t>(S::HALF + j) : 0u;
      threefry2x32(key.k0, key.k1, x0, x1);
      put(j, x0);
      if (S::HALF + j < S::F) put(S::HALF + j, x1);
    }
  }
}

// Emits y[t] for every t in [0, H) through emit(t, value).  `stage` is H floats private to the
// calling thread (its own shared-memory action row: emit may overwrite it, every staged value is
// in a register by then).  `scale` may live in shared or constant memory (F floats).
template <int H, int MODE, typename Emit>
__device__ __forceinline__ void colored_noise_row(Key2 rng, const float* __restrict__ scale, float* stage,
                                                  uint32_t* bits_out, Emit emit) {
  using S = NoiseShape<H>;
  constexpr detail::Twiddle<H> tw{};
  constexpr int LASTK = S::EVEN ? S::F - 1 : S::F;  // bins [1, LASTK) have weight 2
  Key2 key_sr, key_si;
  split3_first2<MODE>(rng, key_sr, key_si);
  stage_normals<H, MODE, false>(key_sr, scale, stage, bits_out);
  stage_normals<H, MODE, true>(key_si, scale, stage, bits_out ? bits_out + S::F : nullptr);

  float sr[S::F], si[S::F];
#pragma unroll
  for (int k = 0; k < S::F; ++k) sr[k] = stage[k];
  si[0] = 0.0f;
#pragma unroll
  for (int k = 1; k < S::F; ++k) si[k] = (k < LASTK) ? stage[S::F + k - 1] : 0.0f;

  // t = 0
  {
    float a = sr[0] * tw.c[0];
#pragma unroll
    for (int k = 1; k < LASTK; ++k) a = fmaf(sr[k], 2.0f * tw.c[0], a);
    if (S::EVEN) a = fmaf(sr[S::F - 1], tw.c[0], a);
    emit(0, a);
  }
#pragma unroll
  for (int t = 1; 2 * t < H; ++t) {
    float a = sr[0] * tw.c[0];
    float b = 0.0f;
#pragma unroll
    for (int k = 1; k < LASTK; ++k) {
      a = fmaf(sr[k], 2.0f * tw.c[(k * t) % H], a);
      b = fmaf(si[k], 2.0f * tw.s[(k * t) % H], b);
    }
    if (S::EVEN) a = fmaf(sr[S::F - 1], (t & 1) ? -tw.c[0] : tw.c[0], a);
    emit(t, a - b);
    emit(H - t, a + b);
  }
  if (S::EVEN) {
    constexpr int t = H / 2;
    float a = sr[0] * tw.c[0];
#pragma unroll
    for (int k = 1; k < LASTK; ++k) a = fmaf(sr[k], (k & 1) ? -2.0f * tw.c[0] : 2.0f * tw.c[0], a);
    a = fmaf(sr[S::F - 1], (t & 1) ? -tw.c[0] : tw.c[0], a);
    emit(t, a);
  }
}

// Host/device helper: the per-bin multiplier table colored_noise_row expects.
//   scale[f] = s_scale[f] / sigma, times sqrt(2) for the DC bin and (even H) the Nyquist bin.
inline void fill_noise_scale(const float* s_scale, float sigma, int horizon, float* scale_out) {
  const int F = horizon / 2 + 1;
  const float inv_sigma = 1.0f / sigma;
  for (int f = 0; f < F; ++f) {
    float s = s_scale[f] * inv_sigma;
    if (f == 0 || (f == F - 1 && horizon % 2 == 0)) s *= 1.41421356f;
    scale_out[f] = s;
  }
}

}  // namespace mbpo
