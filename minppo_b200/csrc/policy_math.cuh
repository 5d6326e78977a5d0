// minppo_b200 -- jax.random.normal restated for the policy sampler (/root/reference/minppo/train.py:158-159):
// bits -> [1,2) mantissa trick -> uniform(nextafter(-1,0), 1) -> sqrt(2) * erf_inv(u), with XLA's single-precision
// erf_inv (Giles' polynomial).  Shared by policy.cu (layer-wise path) and policy_fused.cuh (single-launch path).
#pragma once

#include "common.cuh"
#include "threefry.cuh"

namespace minppo {

// XLA's f32 erf_inv (M. Giles, "Approximating the erfinv function"): w = -log1p(-x^2); two degree-8 polynomials
MINPPO_DEVINL float erfinv_xla(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.f) {
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
    w = sqrtf(w) - 3.f;
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
  return fabsf(x) == 1.f ? copysignf(INFINITY, x) : p * x;
}

// jax.random.normal(key, shape, float32) element from its 32 random bits (jax/_src/random.py: _normal_real, _uniform)
MINPPO_DEVINL float normal_from_bits(uint32_t bits) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;        // [0, 1)
  const float lo = -0.99999994f;                                             // nextafter(-1, 0)
  const float span = __fsub_rn(1.0f, lo);                                    // (maxval - minval) in f32
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, span), lo));
  return __fmul_rn(1.41421356237309504880f, erfinv_xla(u));
}

}  // namespace minppo
