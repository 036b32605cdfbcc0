// Parameter blocks and host entry points of the training kernels (train.cu), shared with the C-ABI layer (capi.cu).
#pragma once
#include <cstdint>
#include "dropout.cuh"

namespace ldot {

struct LnBwdParams {
  const void* dy; long long ld_dy; int dy_f32;
  const void* x; long long ld_x; int x_f32;      // the LayerNorm INPUT (pre-normalisation sum)
  const float* gamma;
  void* dx; long long ld_dx; int dx_f32;
  float* dgamma; float* dbeta; float* dxsum;     // fp32 [H], accumulated (+=); dxsum may be null
  long long rows;
  int fmt;
  // hidden-state dropout of the Linear that fed this LayerNorm, fused (16-bit dy / x / dx only): when dx_masked is
  // given, dx_masked = mask * dx / keep (row pitch ld_dx) is written next to dx and dxsum sums the masked values
  void* dx_masked;
  DropKey drop;
};

struct EmbedImagePreParams {
  const float* lin; const float* box;
  const float* img_g; const float* img_b;
  const float* pos_w; const float* pos_bias;
  const float* pos_g; const float* pos_b;
  const float* type1;
  float* q; float* spre;
  long long rows;
};

struct AdamParams {
  float* p; const float* g; float* m; float* v; uint16_t* p16;
  long long n;
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;   // bc1 = 1 - b1^t, bc2_sqrt = sqrt(1 - b2^t)
  const float* sumsq; float max_norm;
  int fmt;
  const float* hyper;   // device { lr, bc1, bc2_sqrt } overriding the by-value fields (null: by value)
};

int gemm_run(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, const float* bias,
             const void* aux, long long ld_aux, void* out, long long ldo, long long M, int N, long long K, int fmt,
             int epi, int out_f32, int accumulate, void* stream);
int ln_bwd_run(const LnBwdParams& p, int H, void* stream);
int attention_bwd_run(const void* qkv, const long long* mask, const void* ctx, const void* dctx, void* dqkv, int B, int S,
                      int H, int heads, int fmt, void* stream, float drop_p = 0.f, unsigned long long seed = 0, int site = 0);
int dropout_run(const void* x, const void* res, void* out, long long rows, int cols, long long ld, float p,
                unsigned long long seed, int site, int fmt, void* stream);
int gelu_run(const void* x, const void* dy, void* out, long long n, int mode, int fmt, void* stream);
int colsum16_run(const void* in, long long ld, long long rows, int N, float* out, int fmt, void* stream);
int embed_text_sum_run(const long long* ids, const long long* pos_ids, long long pos_batch_stride, const void* word,
                       const void* pos, const void* type0, float* out, int B, int L, int H, int vocab, int max_pos,
                       int fmt, void* stream);
int embed_scatter_run(const float* dx, const long long* ids, const long long* pos_ids, long long pos_batch_stride,
                      float* dword, float* dpos, int B, int L, int H, int vocab, int max_pos, void* stream);
int embed_image_pre_run(const EmbedImagePreParams& p, int H, void* stream);
int pos_wgrad_run(const float* dq, const float* box, long long rows, int H, float* dw, void* stream);
int nll_bwd_run(const float* s, const long long* pos, long long bq, long long bc, const float* upstream, int reduction,
                void* ds, long long ld_ds, int fmt, void* stream);
int sumsq_run(const float* g, long long n, float* out, void* stream);
int adamw_run(const AdamParams& a, void* stream);

}  // namespace ldot
