// Training-step kernels of the two towers (the backward of train_itm.py:252-289 through
// dvl/models/bi_encoder.py's towers).  The dense contractions of the backward (dgrad / wgrad) are the tcgen05 kernel
// of linear_tc.cuh in its MN-major / split-K forms (gemm_ops.cu: gemm_run); this file holds everything else:
//
//   ln_bwd            LayerNorm backward (uniter_model/model/layer.py:108-115,149-156 under autograd): dx, and the
//                     column reductions d gamma, d beta, sum_rows dx (= the bias gradient of the Linear that fed it)
//   attention_bwd     softmax(Q K^T / 8 + mask) V backward per (sequence, head): dQ | dK | dV        (layer.py:80-101)
//   gelu / gelu_bwd   elementwise erf-GELU and its derivative (FFN-up keeps its pre-activation in training)
//   colsum16          fp32 column sums of a 16-bit matrix, accumulated (bias gradients)
//   embed_text_sum / embed_scatter          text-embedding backward (model.py:233-246; word row 0 is padding_idx)
//   embed_image_pre / pos_wgrad             image-embedding backward (model.py:262-273)
//   nll_bwd           d scores of the in-batch NLL (dvl/models/bi_encoder.py:632-640 under autograd)
//   adamw / sumsq     fused decoupled-weight-decay Adam over one flat tensor (bi_encoder.py:566-576 -> AdamW), and
//                     the squared gradient norm for clipping (train_itm.py:262-267)
// Storage formats are runtime flags here (fmt: 0 = fp16, 1 = bf16; *_f32: the tensor is fp32): these kernels are
// bandwidth- or latency-bound, a branch per 16-byte vector is free.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include "host_common.h"
#include "prof.h"
#include "rowops.cuh"
#include "gelu.cuh"
#include "train_params.h"

namespace ldot {

__device__ __forceinline__ void ld8(const void* base, long long idx, int f32, int fmt, float (&f)[8]) {
  if (f32) load8_f32(static_cast<const float*>(base) + idx, f);
  else if (fmt == 1) load8_16<1>(static_cast<const uint16_t*>(base) + idx, f);
  else load8_16<0>(static_cast<const uint16_t*>(base) + idx, f);
}
__device__ __forceinline__ void st8(void* base, long long idx, int f32, int fmt, const float (&f)[8]) {
  if (f32) {
    float4* p = reinterpret_cast<float4*>(static_cast<float*>(base) + idx);
    p[0] = make_float4(f[0], f[1], f[2], f[3]);
    p[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else if (fmt == 1) store8_16<1>(static_cast<uint16_t*>(base) + idx, f);
  else store8_16<0>(static_cast<uint16_t*>(base) + idx, f);
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward

// one warp per row, rows strided over the grid; per-lane column partials live in registers, are combined per block in
// shared memory and leave with one atomicAdd per column per block
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdParams p) {
  constexpr int H = NV * 256;
  extern __shared__ float ln_red[];   // [3][H]
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) ln_red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[NV][8], ab[NV][8], ax[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[v][j] = ab[v][j] = ax[v][j] = 0.f;
  float gam[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) load8_f32(p.gamma + (v * 32 + lane) * 8, gam[v]);

  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < p.rows;
       row += static_cast<long long>(gridDim.x) * 8) {
    float x[NV][8], dy[NV][8];
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * 32 + lane) * 8;
      ld8(p.x, row * p.ld_x + col, p.x_f32, p.fmt, x[v]);
      ld8(p.dy, row * p.ld_dy + col, p.dy_f32, p.fmt, dy[v]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[v][j];
    }
    const float mean = warp_sum(s) / static_cast<float>(H);
    float q = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[v][j] - mean;
        q = fmaf(d, d, q);
      }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(H) + kLnEps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (x[v][j] - mean) * rstd;
        const float g = dy[v][j] * gam[v][j];
        x[v][j] = xh;
        c1 += g;
        c2 = fmaf(g, xh, c2);
      }
    c1 = warp_sum(c1) / static_cast<float>(H);
    c2 = warp_sum(c2) / static_cast<float>(H);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float dx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = dy[v][j] * gam[v][j];
        dx[j] = rstd * (g - c1 - x[v][j] * c2);
        ag[v][j] = fmaf(dy[v][j], x[v][j], ag[v][j]);
        ab[v][j] += dy[v][j];
        ax[v][j] += dx[j];
      }
      st8(p.dx, row * p.ld_dx + (v * 32 + lane) * 8, p.dx_f32, p.fmt, dx);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = (v * 32 + lane) * 8 + j;
      atomicAdd(&ln_red[col], ag[v][j]);
      atomicAdd(&ln_red[H + col], ab[v][j]);
      atomicAdd(&ln_red[2 * H + col], ax[v][j]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(p.dgamma + i, ln_red[i]);
    atomicAdd(p.dbeta + i, ln_red[H + i]);
    if (p.dxsum) atomicAdd(p.dxsum + i, ln_red[2 * H + i]);
  }
}

#define LDOT_NV_DISPATCH(H, CALL)                                   \
  switch ((H) / 256) {                                              \
    case 1: { constexpr int NV = 1; CALL; break; }                  \
    case 2: { constexpr int NV = 2; CALL; break; }                  \
    case 3: { constexpr int NV = 3; CALL; break; }                  \
    case 4: { constexpr int NV = 4; CALL; break; }                  \
    case 6: { constexpr int NV = 6; CALL; break; }                  \
    default: return set_error(kErrArg, "hidden size %d not supported (256 x {1,2,3,4,6})", (H)); \
  }

static unsigned row_grid(long long rows) {
  long long b = (rows + 7) / 8;
  const long long cap = 148 * 4;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

int ln_bwd_run(const LnBwdParams& p, int H, void* stream) {
  LDOT_REQUIRE(p.rows >= 1 && H % 256 == 0, "layernorm_bwd: bad shape rows=%lld H=%d", p.rows, H);
  LDOT_REQUIRE(p.ld_dy % 8 == 0 && p.ld_x % 8 == 0 && p.ld_dx % 8 == 0, "layernorm_bwd: row pitches must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcLayerNorm, st, 0.0, static_cast<double>(p.rows) * H * 6.0);
  const size_t smem = static_cast<size_t>(3) * H * sizeof(float);
  LDOT_NV_DISPATCH(H, (ln_bwd_kernel<NV><<<row_grid(p.rows), 256, smem, st>>>(p)))
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ attention backward
constexpr int kHd = 64;
constexpr int kPitch = 66;   // shared-memory row pitch in 16-bit elements (33 words: conflict-free row-strided reads)

__device__ __forceinline__ float h2f(uint16_t u, int fmt) {
  if (fmt == 1) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}
__device__ __forceinline__ uint16_t f2h(float f, int fmt) {
  if (fmt == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(f);
  return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float dot64(const uint16_t* a, const uint16_t* b, int fmt) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kHd; c += 2) {
    const uint32_t ua = *reinterpret_cast<const uint32_t*>(a + c), ub = *reinterpret_cast<const uint32_t*>(b + c);
    const float2 x = fmt == 1 ? cvt2<1>(ua) : cvt2<0>(ua);
    const float2 y = fmt == 1 ? cvt2<1>(ub) : cvt2<0>(ub);
    s = fmaf(x.x, y.x, s);
    s = fmaf(x.y, y.y, s);
  }
  return s;
}

// grid (heads, B), 256 threads.  qkv [B * S, 3 H] (saved forward activations), ctx / dctx [B * S, H], dqkv [B * S, 3 H].
// The probabilities are recomputed (S <= 128), never stored by the forward.
__global__ void __launch_bounds__(256) attention_bwd_kernel(const uint16_t* __restrict__ qkv, const long long* __restrict__ mask,
                                                            const uint16_t* __restrict__ ctx, const uint16_t* __restrict__ dctx,
                                                            uint16_t* __restrict__ dqkv, int S, int H, int fmt) {
  extern __shared__ __align__(16) uint8_t ab_smem[];
  uint16_t* sQ = reinterpret_cast<uint16_t*>(ab_smem);
  uint16_t* sK = sQ + S * kPitch;
  uint16_t* sV = sK + S * kPitch;
  uint16_t* sdO = sV + S * kPitch;
  float* sP = reinterpret_cast<float*>(sdO + S * kPitch);   // (4 S rows of 132 B: 4-byte aligned)
  float* sD = sP + S * (S + 1);
  float* sMb = sD + S;
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long tok0 = static_cast<long long>(b) * S;
  const int ld = 3 * H;

  for (int i = tid; i < S * 8 * 4; i += blockDim.x) {
    const int mat = i / (S * 8), rem = i - mat * S * 8;
    const int r = rem >> 3, c = (rem & 7) * 8;
    const uint16_t* src = mat < 3 ? qkv + (tok0 + r) * ld + mat * H + head * kHd + c
                                  : dctx + (tok0 + r) * H + head * kHd + c;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
    uint32_t* dst = reinterpret_cast<uint32_t*>(sQ + mat * S * kPitch + r * kPitch + c);
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
  for (int j = tid; j < S; j += blockDim.x) sMb[j] = mask[tok0 + j] != 0 ? 0.f : -10000.f;
  __syncthreads();

  // scores
  for (int idx = tid; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    sP[i * (S + 1) + j] = fmaf(dot64(sQ + i * kPitch, sK + j * kPitch, fmt), 0.125f, sMb[j]);
  }
  __syncthreads();
  // row softmax + D_i = dO_i . O_i
  for (int i = warp; i < S; i += 8) {
    float* row = sP + i * (S + 1);
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float e = __expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < S; j += 32) row[j] *= inv;
    const uint32_t uo = __ldg(reinterpret_cast<const uint32_t*>(ctx + (tok0 + i) * H + head * kHd + lane * 2));
    const uint32_t ud = *reinterpret_cast<const uint32_t*>(sdO + i * kPitch + lane * 2);
    const float2 o2 = fmt == 1 ? cvt2<1>(uo) : cvt2<0>(uo);
    const float2 d2 = fmt == 1 ? cvt2<1>(ud) : cvt2<0>(ud);
    const float d = warp_sum(fmaf(o2.x, d2.x, o2.y * d2.y));
    if (lane == 0) sD[i] = d;
  }
  __syncthreads();
  // dV[j, d] = sum_i P[i, j] dO[i, d]
  for (int idx = tid; idx < S * kHd; idx += blockDim.x) {
    const int j = idx >> 6, d = idx & 63;
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc = fmaf(sP[i * (S + 1) + j], h2f(sdO[i * kPitch + d], fmt), acc);
    dqkv[(tok0 + j) * ld + 2 * H + head * kHd + d] = f2h(acc, fmt);
  }
  __syncthreads();
  // dS = P * (dP - D) / 8 in place, dP[i, j] = dO_i . V_j
  for (int idx = tid; idx < S * S; idx += blockDim.x) {
    const int i = idx / S, j = idx - i * S;
    const float dp = dot64(sdO + i * kPitch, sV + j * kPitch, fmt);
    sP[i * (S + 1) + j] *= (dp - sD[i]) * 0.125f;
  }
  __syncthreads();
  // dQ[i, d] = sum_j dS[i, j] K[j, d];  dK[j, d] = sum_i dS[i, j] Q[i, d]
  for (int idx = tid; idx < S * kHd; idx += blockDim.x) {
    const int r = idx >> 6, d = idx & 63;
    float aq = 0.f, ak = 0.f;
    for (int t = 0; t < S; ++t) {
      aq = fmaf(sP[r * (S + 1) + t], h2f(sK[t * kPitch + d], fmt), aq);
      ak = fmaf(sP[t * (S + 1) + r], h2f(sQ[t * kPitch + d], fmt), ak);
    }
    dqkv[(tok0 + r) * ld + head * kHd + d] = f2h(aq, fmt);
    dqkv[(tok0 + r) * ld + H + head * kHd + d] = f2h(ak, fmt);
  }
}

int attention_bwd_run(const void* qkv, const long long* mask, const void* ctx, const void* dctx, void* dqkv, int B, int S,
                      int H, int heads, int fmt, void* stream) {
  LDOT_REQUIRE(B >= 1 && S >= 1 && S <= 128, "attention_bwd: bad shape B=%d S=%d (S <= 128)", B, S);
  LDOT_REQUIRE(H == heads * kHd, "attention_bwd: hidden %d must be heads (%d) x 64", H, heads);
  LDOT_REQUIRE(B <= 65535, "attention_bwd: batch %d > 65535 (split the batch)", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(4) * S * kPitch * 2 + (static_cast<size_t>(S) * (S + 1) + 2 * S) * 4;
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    LDOT_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  KernelScope ks(kKcAttention, st, 10.0 * B * static_cast<double>(S) * S * H, static_cast<double>(B) * S * H * 16.0);
  attention_bwd_kernel<<<dim3(heads, B), 256, smem, st>>>(static_cast<const uint16_t*>(qkv), mask,
                                                           static_cast<const uint16_t*>(ctx),
                                                           static_cast<const uint16_t*>(dctx),
                                                           static_cast<uint16_t*>(dqkv), S, H, fmt);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ GELU (elementwise)
// mode 0: out = gelu(x);  mode 1: out = dy * gelu'(x)
__global__ void __launch_bounds__(256) gelu_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ dy,
                                                   uint16_t* __restrict__ out, long long n8, int mode, int fmt) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    ld8(x, i * 8, 0, fmt, f);
    if (mode == 1) {
      ld8(dy, i * 8, 0, fmt, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * gelu_erf_grad(f[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
    }
    st8(out, i * 8, 0, fmt, f);
  }
}

static unsigned flat_grid(long long n_vec) {
  long long b = (n_vec + 255) / 256;
  const long long cap = 148 * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

int gelu_run(const void* x, const void* dy, void* out, long long n, int mode, int fmt, void* stream) {
  LDOT_REQUIRE(n >= 0 && n % 8 == 0, "gelu: element count %lld must be a multiple of 8", n);
  LDOT_REQUIRE((mode == 1) == (dy != nullptr), "gelu: dy is required by (and only by) the backward mode");
  if (n == 0) return kOk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcCast, st, 0.0, static_cast<double>(n) * (mode ? 6.0 : 4.0));
  gelu_kernel<<<flat_grid(n / 8), 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(dy),
                                                static_cast<uint16_t*>(out), n / 8, mode, fmt);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ column sums
// out[c] += sum_r in[r, c];  block = 32 column octets x 8 row lanes, grid (ceil(N / 256), row slabs)
__global__ void __launch_bounds__(256) colsum16_kernel(const uint16_t* __restrict__ in, long long ld, long long rows, int N,
                                                       float* __restrict__ out, int fmt) {
  __shared__ float red[8][256];
  const int oct = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + oct) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < N) {
    for (long long r = static_cast<long long>(blockIdx.y) * 8 + ry; r < rows; r += static_cast<long long>(gridDim.y) * 8) {
      float f[8];
      ld8(in, r * ld + col, 0, fmt, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][oct * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  if (blockIdx.x * 256 + c < N) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += red[y][c];
    atomicAdd(out + blockIdx.x * 256 + c, s);
  }
}

int colsum16_run(const void* in, long long ld, long long rows, int N, float* out, int fmt, void* stream) {
  LDOT_REQUIRE(rows >= 1 && N >= 8 && N % 8 == 0 && ld % 8 == 0, "colsum: bad shape rows=%lld N=%d", rows, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long slabs = (rows + 63) / 64;
  slabs = slabs > 64 ? 64 : slabs;
  KernelScope ks(kKcCast, st, 0.0, static_cast<double>(rows) * N * 2.0);
  colsum16_kernel<<<dim3((N + 255) / 256, static_cast<unsigned>(slabs)), 256, 0, st>>>(
      static_cast<const uint16_t*>(in), ld, rows, N, out, fmt);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ text embeddings
// sum[tok, :] = word[ids] + pos[pos_ids] + type0 in fp32 (the LayerNorm input of embed_text_kernel, recomputed)
template <int NV>
__global__ void __launch_bounds__(256) embed_text_sum_kernel(const long long* __restrict__ ids, const long long* __restrict__ pos_ids,
                                                             long long pos_batch_stride, const uint16_t* __restrict__ word,
                                                             const uint16_t* __restrict__ pos, const uint16_t* __restrict__ type0,
                                                             float* __restrict__ out, int B, int L, int vocab, int max_pos,
                                                             int fmt) {
  constexpr int H = NV * 256;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(B) * L) return;
  const int b = static_cast<int>(tok / L), l = static_cast<int>(tok - static_cast<long long>(b) * L);
  long long id = ids[tok];
  long long pid = pos_ids[b * pos_batch_stride + l];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  pid = pid < 0 ? 0 : (pid >= max_pos ? max_pos - 1 : pid);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    float w[8], q[8], t[8];
    ld8(word, id * H + col, 0, fmt, w);
    ld8(pos, pid * H + col, 0, fmt, q);
    ld8(type0, col, 0, fmt, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = w[j] + q[j] + t[j];
    st8(out, tok * H + col, 1, fmt, w);
  }
}

// dword[ids[tok], :] += dx[tok, :] (skipping padding_idx 0, model.py:221-222), dpos[pos_ids[tok], :] += dx[tok, :]
template <int NV>
__global__ void __launch_bounds__(256) embed_scatter_kernel(const float* __restrict__ dx, const long long* __restrict__ ids,
                                                            const long long* __restrict__ pos_ids, long long pos_batch_stride,
                                                            float* __restrict__ dword, float* __restrict__ dpos, int B, int L,
                                                            int vocab, int max_pos) {
  constexpr int H = NV * 256;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(B) * L) return;
  const int b = static_cast<int>(tok / L), l = static_cast<int>(tok - static_cast<long long>(b) * L);
  long long id = ids[tok];
  long long pid = pos_ids[b * pos_batch_stride + l];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  pid = pid < 0 ? 0 : (pid >= max_pos ? max_pos - 1 : pid);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    float f[8];
    load8_f32(dx + tok * H + col, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (id != 0) atomicAdd(dword + id * H + col + j, f[j]);
      atomicAdd(dpos + pid * H + col + j, f[j]);
    }
  }
}

int embed_text_sum_run(const long long* ids, const long long* pos_ids, long long pos_batch_stride, const void* word,
                       const void* pos, const void* type0, float* out, int B, int L, int H, int vocab, int max_pos,
                       int fmt, void* stream) {
  LDOT_REQUIRE(B >= 1 && L >= 1 && H % 256 == 0, "embed_text_sum: bad shape B=%d L=%d H=%d", B, L, H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(B) * L + 7) / 8);
  KernelScope ks(kKcEmbed, st, 0.0, static_cast<double>(B) * L * H * 8.0);
  LDOT_NV_DISPATCH(H, (embed_text_sum_kernel<NV><<<blocks, 256, 0, st>>>(
                          ids, pos_ids, pos_batch_stride, static_cast<const uint16_t*>(word),
                          static_cast<const uint16_t*>(pos), static_cast<const uint16_t*>(type0), out, B, L, vocab,
                          max_pos, fmt)))
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int embed_scatter_run(const float* dx, const long long* ids, const long long* pos_ids, long long pos_batch_stride,
                      float* dword, float* dpos, int B, int L, int H, int vocab, int max_pos, void* stream) {
  LDOT_REQUIRE(B >= 1 && L >= 1 && H % 256 == 0, "embed_scatter: bad shape B=%d L=%d H=%d", B, L, H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(B) * L + 7) / 8);
  KernelScope ks(kKcEmbed, st, 0.0, static_cast<double>(B) * L * H * 12.0);
  LDOT_NV_DISPATCH(H, (embed_scatter_kernel<NV><<<blocks, 256, 0, st>>>(dx, ids, pos_ids, pos_batch_stride, dword, dpos, B,
                                                                         L, vocab, max_pos)))
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ image embeddings
// Recomputes the two intermediate LayerNorm inputs of embed_image_kernel for the backward:
//   q[r, :]    = W_pos box[r] + b_pos                                  (input of pos_layer_norm)
//   spre[r, :] = LN_img(lin[r]) + LN_pos(q[r]) + type1                 (input of img_embeddings.LayerNorm)

template <int NV>
__device__ __forceinline__ void warp_ln_f32(float (&x)[NV][8], int lane, const float* __restrict__ gamma,
                                            const float* __restrict__ beta) {
  constexpr int H = NV * 256;
  float s = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[v][j];
  const float mean = warp_sum(s) / static_cast<float>(H);
  float q = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = x[v][j] - mean;
      q = fmaf(d, d, q);
    }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(H) + kLnEps);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    float g[8], b[8];
    load8_f32(gamma + col, g);
    load8_f32(beta + col, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[v][j] = fmaf((x[v][j] - mean) * rstd, g[j], b[j]);
  }
}

template <int NV>
__global__ void __launch_bounds__(256) embed_image_pre_kernel(const EmbedImagePreParams p) {
  constexpr int H = NV * 256;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= p.rows) return;
  float bx[7];
#pragma unroll
  for (int c = 0; c < 7; ++c) bx[c] = __ldg(p.box + tok * 7 + c);
  float a[NV][8], q[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    load8_f32(p.lin + tok * H + col, a[v]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float* w = p.pos_w + static_cast<long long>(col + j) * 7;
      float s = __ldg(p.pos_bias + col + j);
#pragma unroll
      for (int c = 0; c < 7; ++c) s = fmaf(__ldg(w + c), bx[c], s);
      q[v][j] = s;
    }
    st8(p.q, tok * H + col, 1, 0, q[v]);
  }
  warp_ln_f32<NV>(a, lane, p.img_g, p.img_b);
  warp_ln_f32<NV>(q, lane, p.pos_g, p.pos_b);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    float t[8];
    load8_f32(p.type1 + col, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[v][j] = a[v][j] + q[v][j] + t[j];
    st8(p.spre, tok * H + col, 1, 0, a[v]);
  }
}

int embed_image_pre_run(const EmbedImagePreParams& p, int H, void* stream) {
  LDOT_REQUIRE(p.rows >= 1 && H % 256 == 0, "embed_image_pre: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcEmbed, st, 0.0, static_cast<double>(p.rows) * H * 12.0);
  LDOT_NV_DISPATCH(H, (embed_image_pre_kernel<NV><<<static_cast<unsigned>((p.rows + 7) / 8), 256, 0, st>>>(p)))
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// dW_pos[h, c] += sum_r dq[r, h] box[r, c]   (pos_linear is [H, 7]: too thin for the tensor core)
__global__ void __launch_bounds__(256) pos_wgrad_kernel(const float* __restrict__ dq, const float* __restrict__ box,
                                                        long long rows, int H, float* __restrict__ dw) {
  const int h = blockIdx.x * 256 + threadIdx.x;
  __shared__ float sbox[64][7];
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long r0 = static_cast<long long>(blockIdx.y) * 64; r0 < rows; r0 += static_cast<long long>(gridDim.y) * 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 7; i += 256) {
      const long long r = r0 + i / 7;
      sbox[i / 7][i % 7] = r < rows ? __ldg(box + r * 7 + i % 7) : 0.f;
    }
    __syncthreads();
    if (h < H) {
      const int n = rows - r0 < 64 ? static_cast<int>(rows - r0) : 64;
      for (int i = 0; i < n; ++i) {
        const float g = __ldg(dq + (r0 + i) * H + h);
#pragma unroll
        for (int c = 0; c < 7; ++c) acc[c] = fmaf(g, sbox[i][c], acc[c]);
      }
    }
  }
  if (h < H) {
#pragma unroll
    for (int c = 0; c < 7; ++c) atomicAdd(dw + static_cast<long long>(h) * 7 + c, acc[c]);
  }
}

int pos_wgrad_run(const float* dq, const float* box, long long rows, int H, float* dw, void* stream) {
  LDOT_REQUIRE(rows >= 1 && H >= 1, "pos_wgrad: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long slabs = (rows + 63) / 64;
  slabs = slabs > 32 ? 32 : slabs;
  KernelScope ks(kKcEmbed, st, 14.0 * rows * H, static_cast<double>(rows) * H * 4.0);
  pos_wgrad_kernel<<<dim3((H + 255) / 256, static_cast<unsigned>(slabs)), 256, 0, st>>>(dq, box, rows, H, dw);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ NLL backward
// ds[i, j] = upstream * (softmax_j(s[i, :]) - [j == pos[i]]) * (1 / bq for the mean reduction), 16-bit [bq, ld_ds]
// (columns bc .. ld_ds - 1 are zero-filled: the matrix is an MN-/K-major GEMM operand afterwards)
__global__ void __launch_bounds__(256) nll_bwd_kernel(const float* __restrict__ s, const long long* __restrict__ pos, long long bq,
                                                      long long bc, const float* __restrict__ upstream, float scale,
                                                      uint16_t* __restrict__ ds, long long ld_ds, int fmt) {
  __shared__ float red[32];
  const long long i = blockIdx.x;
  const float* row = s + i * bc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (long long j = threadIdx.x; j < bc; j += 256) mx = fmaxf(mx, row[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (long long j = threadIdx.x; j < bc; j += 256) sum += __expf(row[j] - mx);
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += red[w];
  const float g = __ldg(upstream) * scale, inv = 1.f / sum;
  const long long pi = pos[i];
  for (long long j = threadIdx.x; j < ld_ds; j += 256) {
    float v = 0.f;
    if (j < bc) v = g * (__expf(row[j] - mx) * inv - (j == pi ? 1.f : 0.f));
    ds[i * ld_ds + j] = f2h(v, fmt);
  }
}

int nll_bwd_run(const float* s, const long long* pos, long long bq, long long bc, const float* upstream, int reduction,
                void* ds, long long ld_ds, int fmt, void* stream) {
  LDOT_REQUIRE(bq >= 1 && bc >= 1 && ld_ds >= bc, "nll_bwd: bad shape bq=%lld bc=%lld ld=%lld", bq, bc, ld_ds);
  LDOT_REQUIRE(bq < (1ll << 31), "nll_bwd: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcNll, st, 0.0, static_cast<double>(bq) * bc * 6.0);
  nll_bwd_kernel<<<static_cast<unsigned>(bq), 256, 0, st>>>(s, pos, bq, bc, upstream,
                                                            reduction == 0 ? 1.f / static_cast<float>(bq) : 1.f,
                                                            static_cast<uint16_t*>(ds), ld_ds, fmt);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ optimiser
// sumsq[0] += sum g^2 (fp32 accumulate per block, one atomicAdd per block)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = g[i];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

int sumsq_run(const float* g, long long n, float* out, void* stream) {
  LDOT_REQUIRE(n >= 0 && g && out, "sumsq: bad arguments");
  if (n == 0) return kOk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcOptim, st, 0.0, static_cast<double>(n) * 4.0);
  sumsq_kernel<<<flat_grid((n + 3) / 4), 256, 0, st>>>(g, n, out);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// AdamW step over one flat fp32 tensor (torch.optim.AdamW semantics):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// g is first scaled by clip = min(1, max_norm / (sqrt(*sumsq) + 1e-6)) when sumsq != null (clip_grad_norm_), and
// the updated parameter is also written to the tower's 16-bit inference copy when p16 != null.

__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
  float clip = 1.f;
  if (a.sumsq != nullptr) {
    const float norm = sqrtf(__ldg(a.sumsq));
    clip = fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  const float step = a.lr / a.bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g = a.g[i] * clip;
    float p = a.p[i] * (1.f - a.lr * a.wd);
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    p -= step * m / (sqrtf(v) / a.bc2_sqrt + a.eps);
    a.p[i] = p;
    a.m[i] = m;
    a.v[i] = v;
    if (a.p16 != nullptr) a.p16[i] = f2h(p, a.fmt);
  }
}

int adamw_run(const AdamParams& a, void* stream) {
  LDOT_REQUIRE(a.n >= 0 && a.p && a.g && a.m && a.v, "adamw: null pointer argument");
  if (a.n == 0) return kOk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcOptim, st, 0.0, static_cast<double>(a.n) * 28.0);
  adamw_kernel<<<flat_grid((a.n + 3) / 4), 256, 0, st>>>(a);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

}  // namespace ldot
