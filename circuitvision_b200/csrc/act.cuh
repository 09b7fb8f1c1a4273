// Activation helpers shared by the GEMM epilogue and the CUDA-core kernels: exact-erf GELU evaluated with the
// Abramowitz & Stegun 7.1.26 erf (|error| <= 1.5e-7) on scalar and on packed fp32x2 (FFMA2) arithmetic.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// exact-erf GELU with erf from Abramowitz & Stegun 7.1.26 (|erf error| <= 1.5e-7, far below the bf16 / fp32-sum noise
// of the value it is applied to): 2 MUFU + ~14 FMA-pipe instructions instead of erff's ~40.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-p, ex, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// The same GELU on a PAIR of values with Blackwell's packed fp32x2 arithmetic (FFMA2 / FMUL2): the polynomial, the
// exponent argument and the final blend each issue once for two elements, which is what the epilogue-bound fc1 GEMMs
// need (the epilogue, not the tensor pipe, limits them: ~100 M GELUs per image).
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void up2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void gelu_fast2(float& x0, float& x1) {
  const uint64_t z = pk2(fabsf(x0) * 0.70710678118654752440f, fabsf(x1) * 0.70710678118654752440f);
  float d0, d1, t0, t1;
  up2(fma2(z, pk2(0.3275911f, 0.3275911f), pk2(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = pk2(t0, t1);
  // negated A&S coefficients: p = -(a1 t + ... + a5 t^5)
  uint64_t p = fma2(t, pk2(-1.061405429f, -1.061405429f), pk2(1.453152027f, 1.453152027f));
  p = fma2(p, t, pk2(-1.421413741f, -1.421413741f));
  p = fma2(p, t, pk2(0.284496736f, 0.284496736f));
  p = fma2(p, t, pk2(-0.254829592f, -0.254829592f));
  p = mul2(p, t);
  float e0, e1, ex0, ex1;
  up2(mul2(mul2(z, z), pk2(-1.4426950408889634f, -1.4426950408889634f)), e0, e1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex0) : "f"(e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex1) : "f"(e1));
  float r0, r1;
  up2(fma2(p, pk2(ex0, ex1), pk2(1.0f, 1.0f)), r0, r1);  // |erf| = 1 - poly * exp(-z^2)
  const uint64_t hx = mul2(pk2(x0, x1), pk2(0.5f, 0.5f));
  up2(fma2(hx, pk2(copysignf(r0, x0), copysignf(r1, x1)), hx), x0, x1);
}

// GELU for the epilogue-bound fc1 GEMMs whose output is rounded to 16 bits anyway.  ncu (profiles/r1_gemm_gelu_ncu_summary.md):
// the epilogue warps are ISSUE-bound (472 instructions per 32-column chunk, 336 of them the A&S GELU above); moving the
// exponential from MUFU to the FMA pipe made it slower (more instructions), so this form minimises the instruction count
// instead: GELU(x) = x * sigmoid(x * P(min(x^2, 30.25))) with an odd minimax polynomial in the logit,
// max |error| 2.3e-5 over the whole real line (fit: scripts/fit_gelu_sigmoid.py) — below half an fp16 ulp for every
// |GELU(x)| > 0.05 and 40x below a bf16 ulp there.  13 instructions per pair instead of 21, 4 MUFU per pair as before.
// The coefficients carry the factor -log2(e) so that the MUFU computes 2^(-g) directly.
__device__ __forceinline__ void gelu_sig2(float& x0, float& x1) {
  const uint64_t x = pk2(x0, x1);
  float u0, u1;
  up2(mul2(x, x), u0, u1);
  const uint64_t u = pk2(fminf(u0, 30.25f), fminf(u1, 30.25f));
  uint64_t g = fma2(u, pk2(3.8351773e-05f, 3.8351773e-05f), pk2(5.5078946e-04f, 5.5078946e-04f));
  g = fma2(g, u, pk2(-1.0535588e-01f, -1.0535588e-01f));
  g = fma2(g, u, pk2(-2.3020522e+00f, -2.3020522e+00f));
  float a0, a1, e0, e1, d0, d1, r0, r1;
  up2(mul2(g, x), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  up2(fma2(pk2(e0, e1), pk2(1.0f, 1.0f), pk2(1.0f, 1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  up2(mul2(x, pk2(r0, r1)), x0, x1);
}

// Same logit polynomial, sigmoid evaluated as 0.5 + 0.5 tanh(g / 2) with ONE MUFU (tanh.approx, relative error 2^-11)
// per element: |GELU error| <= |x| * 2.5e-4 * |tanh|, below half an fp16 ulp of the result everywhere.
__device__ __forceinline__ void gelu_tanh2(float& x0, float& x1) {
  const uint64_t x = pk2(x0, x1);
  float u0, u1;
  up2(mul2(x, x), u0, u1);
  const uint64_t u = pk2(fminf(u0, 30.25f), fminf(u1, 30.25f));
  // g / 2 with g = x (c1 + c3 u + c5 u^2 + c7 u^3)
  uint64_t g = fma2(u, pk2(-1.3291712e-05f, -1.3291712e-05f), pk2(-1.9088908e-04f, -1.9088908e-04f));
  g = fma2(g, u, pk2(3.6513564e-02f, 3.6513564e-02f));
  g = fma2(g, u, pk2(7.9783049e-01f, 7.9783049e-01f));
  float a0, a1, t0, t1;
  up2(mul2(g, x), a0, a1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  const uint64_t hx = mul2(x, pk2(0.5f, 0.5f));
  up2(fma2(hx, pk2(t0, t1), hx), x0, x1);
}

// The same tanh-form GELU on a packed fp16 PAIR, for epilogues whose output is fp16 anyway: every step is one half2 instruction
// on a 32-bit register (no 64-bit register pairing, one MUFU.TANH per TWO elements): 10 instructions + 1 MUFU per pair against
// ~19 + 2 of the fp32x2 form as compiled (the pair packing / unpacking around the scalar clamps and MUFUs costs moves).  The logit
// polynomial is evaluated in v = x^2 / 32 <= 0.9453 so that every coefficient is O(1) in fp16:
//   g / 2 = x (0.79783049 + 1.16843405 v - 0.19547042 v^2 - 0.43554282 v^3)
// Error budget: x is rounded to fp16 first (|GELU'| <= 1.13, i.e. about the final rounding again), the fp16 Horner steps add
// ~1.5e-3 relative to the logit (<= 6e-4 absolute after tanh), tanh.approx.f16x2 2^-11 absolute: |GELU error| <~ 4e-4 |x|, about
// 2 fp16 ulps of the result where it matters; the parity fixtures decide (tests/test_sam2_golden_gpu.py).
__device__ __forceinline__ uint32_t h2const(float v) {
  const __half2 h = __float2half2_rn(v);
  return *(const uint32_t*)&h;
}
__device__ __forceinline__ uint32_t gelu_tanh_h2(uint32_t x) {
  uint32_t xs, v, g, a, t, hx, y;
  asm("mul.f16x2 %0, %1, %2;" : "=r"(xs) : "r"(x), "r"(h2const(0.03125f)));
  asm("mul.f16x2 %0, %1, %2;" : "=r"(v) : "r"(xs), "r"(x));
  asm("min.f16x2 %0, %1, %2;" : "=r"(v) : "r"(v), "r"(h2const(0.9453125f)));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(g) : "r"(v), "r"(h2const(-0.43554282f)), "r"(h2const(-0.19547042f)));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(g) : "r"(g), "r"(v), "r"(h2const(1.16843405f)));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(g) : "r"(g), "r"(v), "r"(h2const(0.79783049f)));
  asm("mul.f16x2 %0, %1, %2;" : "=r"(a) : "r"(g), "r"(x));
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(a));
  asm("mul.f16x2 %0, %1, %2;" : "=r"(hx) : "r"(x), "r"(h2const(0.5f)));
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(y) : "r"(hx), "r"(t), "r"(hx));
  return y;
}

}  // namespace cvb
