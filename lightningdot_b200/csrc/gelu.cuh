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

// d/dx of the erf-GELU above, from the same logit polynomial: gelu'(x) = Phi(x) + x phi(x), Phi = sigmoid(x P(x^2)),
// phi(x) = exp(-x^2 / 2) / sqrt(2 pi).  Used by the FFN-up backward epilogue (ACT == 2).
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float x2 = x * x;
  float p = fmaf(-3.2289885893987957e-06f, x2, 8.823812822811306e-05f);
  p = fmaf(p, x2, 0.00036027454189024866f);
  p = fmaf(p, x2, -0.10522668808698654f);
  p = fmaf(p, x2, -2.3020453453063965f);
  float e, r, g;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * p));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(x2 * -0.7213475204444817f));
  return fmaf(x * 0.3989422804014327f, g, r);
}

}  // namespace ldot
