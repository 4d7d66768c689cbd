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
__device__ __forceinline__ float erf_inv_f32(float x) {
  float w = -log1pf(-(x * x));
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
  return (fabsf(x) == 1.0f) ? __int_as_float(0x7F800000) * x : p * x;
}

// jax.random.normal for one 32-bit word: sqrt(2) * erf_inv(uniform(nextafter(-1,0), 1)).
// hi - lo = 1 - (-0.99999994) rounds to exactly 2.0f in float32.
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = -0.99999994f;  // nextafter(-1, 0)
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float u = fmaxf(lo, f * 2.0f + lo);
  return 1.41421356f * erf_inv_f32(u);
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

}  // namespace mbpo
