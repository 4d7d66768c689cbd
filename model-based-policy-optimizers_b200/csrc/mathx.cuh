// bits -> uniform -> normal, as jax.random does it (jax/_src/random.py _uniform,
// _normal_real) with XLA's float32 erf_inv (Giles' polynomials, xla ErfInv32).
// Call sites in the reference: mbpo/utils/general_utils.py:190-191.
#pragma once
#include <stdint.h>

namespace mbpo {

// uniform in [lo, hi): mantissa trick, then max(lo, f*(hi-lo)+lo)
__device__ __forceinline__ float bits_to_uniform(uint32_t bits, float lo, float hi) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  return fmaxf(lo, __fadd_rn(__fmul_rn(f, hi - lo), lo));  // unfused, as the oracle
}

// XLA ErfInv32: w = -log1p(-x*x); w < 5 ? poly(w-2.5) : poly(sqrt(w)-3); result p*x.
// Like the reference formula, x*x is rounded to float32 first (this rounding dominates the
// tail: for |x| -> 1 it is a several-percent perturbation of 1 - x*x, so a "more accurate"
// (1-x)(1+x) would NOT match JAX).  log1p(-t) is then evaluated as log(1 - t) with the hardware
// log2 (MUFU.LG2): 1 - t is exact for t >= 0.5 (Sterbenz) and otherwise perturbs w by <= 6e-8
// absolute, i.e. the result by <= ~2e-8 relative (|p'(w)| <= 0.25) -- below the 1-2 ulp spread
// between log1p implementations (XLA's own is a polynomial), ~20 instructions cheaper.
// GUARD = false drops the |x| == 1 -> +-inf special case for callers whose argument is provably
// inside (-1, 1).
// log(v) = lg2(v) * ln 2 with the bare MUFU.LG2 (lg2.approx.ftz): what __logf computes for a normal v, without
// the denormal-input guard nvcc wraps around it (three instructions per call).  The callers' arguments are 0
// (-> -inf in both) or >= 2^-24.
__device__ __forceinline__ float log_normal_arg(float v) {
  float l;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(v));
  return l * 0.693147182f;
}

template <bool GUARD = true>
__device__ __forceinline__ float erf_inv_f32(float x) {
  float w = -log_normal_arg(1.0f - __fmul_rn(x, x));
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, w, 3.43273939e-07f);
    p = fmaf(p, w, -3.5233877e-06f);
    p = fmaf(p, w, -4.39150654e-06f);
    p = fmaf(p, w, 0.00021858087f);
    p = fmaf(p, w, -0.00125372503f);
    p = fmaf(p, w, -0.00417768164f);
    p = fmaf(p, w, 0.246640727f);
    p = fmaf(p, w, 1.50140941f);
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, w, 0.000100950558f);
    p = fmaf(p, w, 0.00134934322f);
    p = fmaf(p, w, -0.00367342844f);
    p = fmaf(p, w, 0.00573950773f);
    p = fmaf(p, w, -0.0076224613f);
    p = fmaf(p, w, 0.00943887047f);
    p = fmaf(p, w, 1.00167406f);
    p = fmaf(p, w, 2.83297682f);
  }
  return GUARD ? ((fabsf(x) == 1.0f) ? __int_as_float(0x7F800000) * x : p * x) : p * x;
}

// jax.random.normal for one 32-bit word: sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1)).
// hi - lo = 1 - (-0.99999994) rounds to exactly 2.0f in float32, so f * 2 is exact and
// u = f * 2 + lo lies in [lo, 0.99999982]: jax's max(lo, u) is the identity and |u| < 1.
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float u = f * 2.0f + lo;
  return 1.41421356f * erf_inv_f32<false>(u);
}

// Sort key of jnp.argsort (stable, ascending; icem_optimizer.py:199): monotone uint32 image
// of the float under JAX's sort order (jax/_src/lax/lax.py _float_to_int_for_sort): the
// IEEE total order with -0.0 == +0.0 and every NaN equal and last.
__device__ __forceinline__ uint32_t total_order_key(float v) {
  uint32_t b = __float_as_uint(v);
  if (v == 0.0f) b = 0u;
  if (v != v) b = 0x7FC00000u;
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ---- bounded-range trigonometry for the pendulum (|angle| <= ~4 rad) ------------------------
// Branch-free float32 sin/cos/atan2 with 1-2 ulp accuracy (the same accuracy class as libm /
// XLA's own polynomial implementations; the parity bar is rel 1e-5).  No large-argument
// slow path: the callers' angles are bounded by pi + max_speed*dt.

// sin and cos of x, |x| < 2^20.  Cody-Waite reduction by pi/2 in three terms, Cephes minimax
// polynomials on [-pi/4, pi/4].
__device__ __forceinline__ void sincos_bounded(float x, float& s, float& c) {
  const float m = fmaf(x, 0.636619747f, 12582912.0f);  // round(x * 2/pi) in the low mantissa bits
  const int q = __float_as_int(m);
  const float fq = m - 12582912.0f;
  float r = fmaf(fq, -1.57079601e+00f, x);
  r = fmaf(fq, -3.13916473e-07f, r);
  r = fmaf(fq, -5.39030253e-15f, r);
  const float r2 = r * r;
  float sp = fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  const float sr = fmaf(sp * r2, r, r);
  float cp = fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  const float cr = fmaf(cp * r2, r2, fmaf(-0.5f, r2, 1.0f));
  const bool odd = q & 1;
  const float s0 = odd ? cr : sr;
  const float c0 = odd ? sr : cr;
  s = __int_as_float(__float_as_int(s0) ^ ((q & 2) << 30));
  c = __int_as_float(__float_as_int(c0) ^ (((q + 1) & 2) << 30));
}

__device__ __forceinline__ float sin_bounded(float x) {
  float s, c;
  sincos_bounded(x, s, c);
  return s;
}

// atan2(y, x) for finite inputs: atan(min/max) by a degree-17 odd minimax polynomial
// (max abs error 7e-8 on [0, 1]) and quadrant fix-ups; atan2(+-0, +-0) = +-0 like NumPy/XLA
// for (0, +0).
// UNIT = true: the caller guarantees max(|x|, |y|) is a normal float (a [cos, sin] pair out of sincos_bounded has
// max >= 0.707): min * MUFU.RCP(max) without the zero test and the denormal scaling __fdividef carries -- the
// same bits as the guarded form for every such input, eight instructions shorter.
template <bool UNIT = false>
__device__ __forceinline__ float atan2_bounded(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float t;
  if (UNIT) {
    float rcp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(mx));
    t = __fmul_rn(rcp, mn);
  } else {
    t = (mx > 0.0f) ? __fdividef(mn, mx) : 0.0f;
  }
  const float u = t * t;
  float q = 2.974586547e-03f;
  q = fmaf(q, u, -1.658116880e-02f);
  q = fmaf(q, u, 4.355351503e-02f);
  q = fmaf(q, u, -7.580576113e-02f);
  q = fmaf(q, u, 1.067893936e-01f);
  q = fmaf(q, u, -1.421420891e-01f);
  q = fmaf(q, u, 1.999413718e-01f);
  q = fmaf(q, u, -3.333316696e-01f);
  float r = fmaf(t * u, q, t);
  if (ay > ax) r = 1.57079637f - r;
  if (x < 0.0f) r = 3.14159274f - r;
  return copysignf(r, y);
}

}  // namespace mbpo
