// In-batch-negative NLL of the bi-encoder (dvl/models/bi_encoder.py:615-656 BiEncoderNllLoss.calc, wrapped by
// dvl/utils.py:114-169 _calc_loss; symmetric use train_itm.py:203-222).
//
//   scores = Q . C^T                                    (dot_product_scores, bi_encoder.py:54-68)
//   scores = (1 - w) scores + w (Q . Cap^T)             (caption mix, bi_encoder.py:625-627)
//   loss   = reduce_i( logsumexp_j scores[i, j] - scores[i, pos_i] );   correct = #{ argmax_j scores[i, j] == pos_i }
//
// The reference computes the score matrix in fp32 (CPU) / returns it to the caller, so it is materialised here too.
// To keep fp32-grade scores on the 16-bit tensor cores, each fp32 operand is split into fp16 hi + lo parts
// (x = hi + lo + O(2^-22 |x|)) and the three significant products are folded into ONE tcgen05 GEMM by
// concatenating along K:  [q_hi | q_lo | q_hi] . [c_hi | c_hi | c_lo]^T  =  q_hi c_hi + q_lo c_hi + q_hi c_lo.
#include <cuda_fp16.h>
#include <cfloat>
#include <cmath>
#include "host_common.h"
#include "prof.h"

namespace ldot {

// side 0: [hi | lo | hi]   side 1: [hi | hi | lo]        (row-major [rows, 3K] fp16)
__global__ void __launch_bounds__(256) split16_kernel(const float* __restrict__ in, long long rows, int K, int side,
                                                      __half* __restrict__ out) {
  const long long total = rows * K;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / K;
    const int c = static_cast<int>(i - r * K);
    const float v = in[i];
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    __half* o = out + r * 3 * K;
    o[c] = hi;
    o[K + c] = side == 0 ? lo : hi;
    o[2 * K + c] = side == 0 ? hi : lo;
  }
}

__device__ __forceinline__ float block_reduce_max(float v, float* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) r = fmaxf(r, sh[w]);
  return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < (blockDim.x >> 5); ++w) r += sh[w];
  return r;
}

// one block per query row
__global__ void __launch_bounds__(256) nll_rows_kernel(const float* __restrict__ s1, const float* __restrict__ s2, float w,
                                                       const long long* __restrict__ pos, long long bc,
                                                       float* __restrict__ s_out, float* __restrict__ row_loss,
                                                       int* __restrict__ row_correct) {
  __shared__ float shf[8];
  __shared__ double shd[8];
  __shared__ unsigned long long shbest[8];
  const long long row = blockIdx.x;
  const float* a = s1 + row * bc;
  const float* b = s2 ? s2 + row * bc : nullptr;
  float* o = s_out + row * bc;
  float mx = -INFINITY;
  // argmax with "first maximal index" semantics: maximise (key(score), ~col)
  unsigned long long best = 0ull;
  for (long long j = threadIdx.x; j < bc; j += blockDim.x) {
    float v = a[j];
    if (b) v = (1.0f - w) * v + w * b[j];
    o[j] = v;
    mx = fmaxf(mx, v);
    uint32_t u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    const unsigned long long cand = (static_cast<unsigned long long>(u) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(j));
    best = cand > best ? cand : best;
  }
  mx = block_reduce_max(mx, shf);
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xFFFFFFFFu, best, off);
    best = t > best ? t : best;
  }
  if ((threadIdx.x & 31) == 0) shbest[threadIdx.x >> 5] = best;
  double sum = 0.0;
  for (long long j = threadIdx.x; j < bc; j += blockDim.x) sum += static_cast<double>(expf(o[j] - mx));
  sum = block_reduce_sum(sum, shd);  // (its barriers also publish shbest)
  if (threadIdx.x == 0) {
    for (int wv = 1; wv < (blockDim.x >> 5); ++wv) best = shbest[wv] > best ? shbest[wv] : best;
    const long long p = pos[row];
    const float lse = mx + static_cast<float>(log(sum));
    row_loss[row] = lse - o[p];
    const long long arg = static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(best & 0xFFFFFFFFu));
    row_correct[row] = arg == p ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) nll_finalize_kernel(const float* __restrict__ row_loss, const int* __restrict__ row_correct,
                                                           long long bq, int reduction, float* __restrict__ loss,
                                                           long long* __restrict__ correct) {
  __shared__ double shd[8];
  double s = 0.0, c = 0.0;
  for (long long i = threadIdx.x; i < bq; i += blockDim.x) {
    s += static_cast<double>(row_loss[i]);
    c += static_cast<double>(row_correct[i]);
  }
  s = block_reduce_sum(s, shd);
  c = block_reduce_sum(c, shd);
  if (threadIdx.x == 0) {
    *loss = static_cast<float>(reduction == 0 ? s / static_cast<double>(bq) : s);
    *correct = static_cast<long long>(c + 0.5);
  }
}

int split16_run(const float* in, long long rows, int K, int side, void* out, void* stream) {
  LDOT_REQUIRE(rows >= 1 && K >= 1 && (side == 0 || side == 1), "split16: bad arguments");
  long long blocks = (rows * K + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  {
    KernelScope ks(kKcCast, static_cast<cudaStream_t>(stream), 0.0, static_cast<double>(rows) * K * 10.0);
    split16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, rows, K, side,
                                                                                                 static_cast<__half*>(out));
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int nll_run(const float* s1, const float* s2, float w, const long long* pos, long long bq, long long bc, int reduction,
            float* s_out, float* row_loss, int* row_correct, float* loss, long long* correct, void* stream) {
  LDOT_REQUIRE(bq >= 1 && bc >= 1 && bc < (1ll << 31) && bq < (1ll << 31), "nll: bad shape %lld x %lld", bq, bc);
  LDOT_REQUIRE(reduction == 0 || reduction == 1, "nll: reduction must be 0 (mean) or 1 (sum)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    KernelScope ks(kKcNll, st, 0.0, static_cast<double>(bq) * bc * (s2 ? 12.0 : 8.0));
    nll_rows_kernel<<<static_cast<unsigned>(bq), 256, 0, st>>>(s1, s2, w, pos, bc, s_out, row_loss, row_correct);
  }
  LDOT_CHECK_LAUNCH();
  {
    KernelScope ks(kKcNll, st);
    nll_finalize_kernel<<<1, 256, 0, st>>>(row_loss, row_correct, bq, reduction, loss, correct);
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

}  // namespace ldot
