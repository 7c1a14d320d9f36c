// Exact-erf GeLU for the GEMM epilogues, written on PAIRS of values with Blackwell's packed fp32
// instructions (fma/mul/add .f32x2: two IEEE fp32 operations per issued instruction).
//
// The h -> 4h epilogue is instruction-issue bound, not latency bound: ncu counts 24.9 k warp
// instructions per 128x256 tile against 4 x 6144 issue slots while the tile's MMAs run (tensor pipe
// 54 % busy).  Everything that is not a MUFU, an abs or a max is therefore issued once per pair here:
// about 10 instructions per element instead of 17.
//
//   GeLU(x) = x * Phi(x) = max(x, 0) - |x| * q,   q = 1 - Phi(|x|) = 0.5 * erfc(|x| / sqrt 2)
//   erfc(z) ~ (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5) exp(-z^2),  t = 1 / (1 + p z)
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 — far below the 16-bit output rounding; the reference
// computes F.gelu's erf form, megatron/model/transformer.py:80,99-104.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emdr2 {

__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t f2_splat(float v) { return f2_pack(v, v); }

// (GeLU(x0), GeLU(x1)).
__device__ __forceinline__ void gelu_erf_pair(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const uint64_t ax = f2_pack(a0, a1);
  const uint64_t xx = f2_pack(x0, x1);
  // t = 1 / (1 + (p / sqrt 2) |x|)
  const uint64_t den = f2_fma(ax, f2_splat(0.3275911f * 0.70710678118654752f), f2_splat(1.0f));
  float d0, d1, t0, t1;
  f2_unpack(den, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = f2_pack(t0, t1);
  // -0.5 * (a1 t + ... + a5 t^5): the sign and the 0.5 are folded into the coefficients
  uint64_t poly = f2_fma(t, f2_splat(-0.5f * 1.061405429f), f2_splat(-0.5f * -1.453152027f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * 1.421413741f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * -0.284496736f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * 0.254829592f));
  poly = f2_mul(poly, t);
  // exp(-x^2 / 2) = 2^(x^2 * -log2(e) / 2)
  const uint64_t arg = f2_mul(f2_mul(xx, xx), f2_splat(-0.5f * 1.4426950408889634f));
  float g0, g1, e0, e1;
  f2_unpack(arg, g0, g1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t nq = f2_mul(poly, f2_pack(e0, e1));                 // -q
  const uint64_t out = f2_fma(ax, nq, f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
  f2_unpack(out, x0, x1);
}

// (d0 * GeLU'(u0), d1 * GeLU'(u1)): the backward of the h -> 4h activation, applied to the gradient pair in place.
//   GeLU'(u) = Phi(u) + u phi(u) = 0.5 + copysign(0.5 - q, u) + u exp(-u^2 / 2) / sqrt(2 pi),   q as above.
// Same approximation and packed arithmetic as the forward (about 11 issued instructions per element, two of them
// MUFU): the scalar version cost ~20 and made dU = (dA . W2) * GeLU'(u) epilogue-bound even on sixteen warps.
__device__ __forceinline__ void gelu_erf_grad_pair(float u0, float u1, float& d0, float& d1) {
  const float a0 = fabsf(u0), a1 = fabsf(u1);
  const uint64_t ax = f2_pack(a0, a1);
  const uint64_t uu = f2_pack(u0, u1);
  const uint64_t den = f2_fma(ax, f2_splat(0.3275911f * 0.70710678118654752f), f2_splat(1.0f));
  float n0, n1, t0, t1;
  f2_unpack(den, n0, n1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(n0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(n1));
  const uint64_t t = f2_pack(t0, t1);
  uint64_t poly = f2_fma(t, f2_splat(-0.5f * 1.061405429f), f2_splat(-0.5f * -1.453152027f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * 1.421413741f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * -0.284496736f));
  poly = f2_fma(poly, t, f2_splat(-0.5f * 0.254829592f));
  poly = f2_mul(poly, t);
  const uint64_t arg = f2_mul(f2_mul(uu, uu), f2_splat(-0.5f * 1.4426950408889634f));
  float g0, g1, e0, e1;
  f2_unpack(arg, g0, g1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(g1));
  const uint64_t ee = f2_pack(e0, e1);                                // exp(-u^2 / 2)
  float h0, h1;
  f2_unpack(f2_fma(poly, ee, f2_splat(0.5f)), h0, h1);                // 0.5 - q  (>= 0)
  h0 = __uint_as_float(__float_as_uint(h0) ^ (__float_as_uint(u0) & 0x80000000u));   // copysign(., u)
  h1 = __uint_as_float(__float_as_uint(h1) ^ (__float_as_uint(u1) & 0x80000000u));
  const uint64_t cdf = f2_add(f2_pack(h0, h1), f2_splat(0.5f));       // Phi(u)
  const uint64_t grad = f2_fma(f2_mul(uu, f2_splat(0.3989422804014327f)), ee, cdf);
  f2_unpack(f2_mul(f2_pack(d0, d1), grad), d0, d1);
}

}  // namespace emdr2
