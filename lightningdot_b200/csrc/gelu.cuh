// erf-GELU and its derivative as used by the tower epilogues and the training kernels.
#pragma once
#include <cuda_runtime.h>

namespace ldot {

// erf-GELU  x * Phi(x) = 0.5 x (1 + erf(x / sqrt 2))  (uniter_model/model/layer.py:31-37) evaluated as
//   x * sigmoid(x * P(x^2)),  P a degree-4 minimax polynomial of the exact logit  ln(Phi / (1 - Phi)) / x:
// max |error| 3.4e-6 over the whole real line (checked against the fp64 erf form in tests/), i.e. ~1 % of one
// fp16 ulp and 0.1 % of one bf16 ulp of the 16-bit output.  8 FP32 instructions + 2 MUFU per element instead of the
// ~23 of an erff()-based form: the FFN-up epilogue is ALU-paced.  The coefficients carry the -log2(e) of the
// sigmoid's exp2.
__device__ __forceinline__ float gelu_erf(float x) {
  const float x2 = x * x;
  float p = fmaf(-3.2289885893987957e-06f, x2, 8.823812822811306e-05f);
  p = fmaf(p, x2, 0.00036027454189024866f);
  p = fmaf(p, x2, -0.10522668808698654f);
  p = fmaf(p, x2, -2.3020453453063965f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// GELU and its derivative together, from the SAME logit polynomial: with u(x) = x P(x^2) (log2 units) and
// s = 1 / (1 + 2^u),  gelu = x s  and  gelu' = s + x ds/dx = s - ln2 x s (1 - s) (P + 2 x^2 P'(x^2)).
// Two MUFU (ex2, rcp) for both values - the former derivative evaluated phi(x) with a third one, which made the GELU'
// epilogue of the FFN-up dgrad MUFU-bound (3 x 32768 per tile / 16 per clock = 6144 cycles against 4608 of MMA).
// max |error| of gelu' against Phi(x) + x phi(x): 1.5e-5 (fp64 and fp32 evaluation, |x| <= 12).  x is clamped to +-20 for
// the polynomial (both functions are saturated far earlier) so that s (1 - s) = 0 never meets an overflowed polynomial.
__device__ __forceinline__ void gelu_erf_both(float x, float& g, float& gp) {
  const float xc = fminf(fmaxf(x, -20.f), 20.f);
  const float x2 = xc * xc;
  float p = fmaf(-3.2289885893987957e-06f, x2, 8.823812822811306e-05f);
  p = fmaf(p, x2, 0.00036027454189024866f);
  p = fmaf(p, x2, -0.10522668808698654f);
  p = fmaf(p, x2, -2.3020453453063965f);
  float dp = fmaf(4.f * -3.2289885893987957e-06f, x2, 3.f * 8.823812822811306e-05f);
  dp = fmaf(dp, x2, 2.f * 0.00036027454189024866f);
  dp = fmaf(dp, x2, -0.10522668808698654f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xc * p));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  g = x * r;
  const float t = fmaf(2.f * x2, dp, p);
  const float sr = fmaf(-r, r, r);
  gp = fmaf(-0.6931471805599453f * xc * sr, t, r);
}

// d/dx of the erf-GELU above (FFN-up backward: ldot_gelu_bwd, and the ACT == 2 epilogue of ldot_gemm)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float g, gp;
  gelu_erf_both(x, g, gp);
  return gp;
}

}  // namespace ldot
