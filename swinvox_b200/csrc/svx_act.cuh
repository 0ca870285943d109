// svx_act.cuh -- activation math shared by the contraction kernels (svx_gemm.cu, svx_mlp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace svx {

// exact-erf GELU (nn.GELU()) with a branch-free erf: Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7, ~14 instructions
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
  const float erf_abs = fmaf(-poly * t, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// The same GELU on two values at once with the packed fp32x2 pipe of sm_100 (FFMA2 / FMUL2): the erf-GELU epilogues are
// bound by instruction issue in the epilogue warps (profiles/r1_ncu_hot_lines_v17.txt), packing halves the FMA-pipe
// instructions per element; the two MUFU ops (rcp, ex2) per element stay scalar.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const uint64_t x = pack2(x0, x1);
  const uint64_t ax = x & 0x7fffffff7fffffffull;
  const uint64_t z = mul2(ax, pack2(0.70710678118654752440f, 0.70710678118654752440f));
  // sqrt(log2 e) * |x| / sqrt 2: its square is z^2 * log2(e), the exponent of exp(-z^2) in base 2
  const uint64_t zs = mul2(ax, pack2(0.84932180028801904272f, 0.84932180028801904272f));
  const uint64_t den = fma2(pack2(0.3275911f, 0.3275911f), z, pack2(1.f, 1.f));
  float d0, d1, t0, t1;
  unpack2(den, d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const uint64_t t = pack2(t0, t1);
  // Horner with NEGATED coefficients: npoly = -(a1 + t(a2 + t(a3 + t(a4 + t a5))))
  uint64_t np = fma2(pack2(-1.061405429f, -1.061405429f), t, pack2(1.453152027f, 1.453152027f));
  np = fma2(np, t, pack2(-1.421413741f, -1.421413741f));
  np = fma2(np, t, pack2(0.284496736f, 0.284496736f));
  np = fma2(np, t, pack2(-0.254829592f, -0.254829592f));
  const uint64_t zz = mul2(zs, zs);
  float q0, q1, e0, e1;
  unpack2(zz, q0, q1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-q1));
  const uint64_t erf_abs = fma2(mul2(np, t), pack2(e0, e1), pack2(1.f, 1.f));     // 1 - poly*t*exp(-z^2)
  const uint64_t s = erf_abs | (x & 0x8000000080000000ull);                       // copysign (erf_abs >= 0)
  const uint64_t hx = mul2(x, pack2(0.5f, 0.5f));
  unpack2(fma2(hx, s, hx), x0, x1);
}

// erf-GELU with ONE MUFU op per element: Abramowitz-Stegun 7.1.28, erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16 (|error| <= 3e-7;
// 8e-7 on the GELU in fp32 arithmetic, three orders below the TF32 rounding of whatever consumes it), written as
//   gelu(x) = relu(x) - 0.5 |x| / P(|x|)^16,   P's coefficients rescaled by 2^(-k/2) so that it takes |x| directly.
// The erf-GELU epilogues are bound by the MUFU pipe (16 lanes per clock and SM) together with instruction issue; this form
// halves the MUFU work of gelu_erf2 (rcp + ex2) at the same FMA-pipe instruction count.
__device__ __forceinline__ void gelu_erf2_fast(float& x0, float& x1) {
  const uint64_t x = pack2(x0, x1);
  const uint64_t ax = x & 0x7fffffff7fffffffull;
  // P scaled by s = 2^(1/16), so that P^16 comes out doubled and its reciprocal is already 0.5 / P^16:
  //   gelu(x) = max(x, 0) - |x| * (0.5 / P(|x|)^16)          (6 FMA + 4 MUL + 1 FMA on the FP32 pipe per element)
  constexpr float s = 1.04427378242741384032f;
  constexpr float c0 = s, c1 = s * 0.0705230784f * 0.70710678118654752440f, c2 = s * 0.0422820123f * 0.5f,
                  c3 = s * 0.0092705272f * 0.35355339059327376220f, c4 = s * 0.0001520143f * 0.25f,
                  c5 = s * 0.0002765672f * 0.17677669529663688110f, c6 = s * 0.0000430638f * 0.125f;
  uint64_t q = fma2(pack2(c6, c6), ax, pack2(c5, c5));
  q = fma2(q, ax, pack2(c4, c4));
  q = fma2(q, ax, pack2(c3, c3));
  q = fma2(q, ax, pack2(c2, c2));
  q = fma2(q, ax, pack2(c1, c1));
  q = fma2(q, ax, pack2(c0, c0));
  q = mul2(q, q);
  q = mul2(q, q);
  q = mul2(q, q);
  q = mul2(q, q);                                      // 2 P^16 (overflows to +inf for |x| > ~14: 1/inf = 0, erf = 1)
  float q0, q1, r0, r1;
  unpack2(q, q0, q1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(q0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(q1));
  unpack2(fma2(ax, pack2(-r0, -r1), pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), x0, x1);
}

}  // namespace svx
