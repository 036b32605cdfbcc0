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
#include "mma_sync.cuh"
#include "dropout.cuh"
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

// one warp per row, rows strided over the grid; per-lane column partials live in registers, are laid side by side in
// shared memory ([warp][3 H], plain stores), summed over the warps column-wise, and leave with one atomicAdd per
// column per block
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const LnBwdParams p) {
  constexpr int H = NV * 256;
  extern __shared__ float ln_red[];   // [8 warps][3][H]
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
  float* mine = ln_red + warp * 3 * H;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    *reinterpret_cast<float4*>(mine + col) = make_float4(ag[v][0], ag[v][1], ag[v][2], ag[v][3]);
    *reinterpret_cast<float4*>(mine + col + 4) = make_float4(ag[v][4], ag[v][5], ag[v][6], ag[v][7]);
    *reinterpret_cast<float4*>(mine + H + col) = make_float4(ab[v][0], ab[v][1], ab[v][2], ab[v][3]);
    *reinterpret_cast<float4*>(mine + H + col + 4) = make_float4(ab[v][4], ab[v][5], ab[v][6], ab[v][7]);
    *reinterpret_cast<float4*>(mine + 2 * H + col) = make_float4(ax[v][0], ax[v][1], ax[v][2], ax[v][3]);
    *reinterpret_cast<float4*>(mine + 2 * H + col + 4) = make_float4(ax[v][4], ax[v][5], ax[v][6], ax[v][7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ln_red[w * 3 * H + i];
    if (i < H) atomicAdd(p.dgamma + i, t);
    else if (i < 2 * H) atomicAdd(p.dbeta + i - H, t);
    else if (p.dxsum) atomicAdd(p.dxsum + i - 2 * H, t);
  }
}

// The hot form (x and dy 16-bit: every encoder-layer LayerNorm of the step): the row stays in registers as the RAW 16-byte
// vectors (re-converted per pass), two of the three column partials live in registers and the third (sum of dx) in the
// warp's own shared-memory plane, so that the kernel fits 128 registers and two blocks (16 rows in flight) per SM - the
// general kernel above needs 226 registers, one block per SM, and ran at a quarter of the HBM rate.
// Hidden-state dropout fused in (p.dx_masked != null): the Linear that fed this LayerNorm was followed by dropout
// (layer.py:111-115: LN(dropout(dense(x)) + residual)), so the residual branch gets dx and the dense branch gets
// mask * dx / keep - written as a second output here, and p.dxsum (the dense bias gradient) sums the MASKED values;
// one kernel instead of LayerNorm backward + dropout + column sums.
template <int NV, int FMT>
__global__ void __launch_bounds__(256, 2) ln_bwd16_kernel(const LnBwdParams p) {
  constexpr int H = NV * 256;
  extern __shared__ float ln_red[];   // [8 warps][3][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ag[NV][8], ab[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[v][j] = ab[v][j] = 0.f;
  float* mine = ln_red + warp * 3 * H;
  float* axs = mine + 2 * H;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    *reinterpret_cast<float4*>(axs + col) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(axs + col + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const DropKey drop = drop_resolve(p.drop);
  const bool masked = p.dx_masked != nullptr && drop.thr != 0;
  const uint16_t* xg = static_cast<const uint16_t*>(p.x);
  const uint16_t* dyg = static_cast<const uint16_t*>(p.dy);
  constexpr float kInvH = 1.f / static_cast<float>(H);

  auto unpack = [](const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 x = cvt2<FMT>(w[t]);
      f[2 * t] = x.x;
      f[2 * t + 1] = x.y;
    }
  };

  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < p.rows;
       row += static_cast<long long>(gridDim.x) * 8) {
    uint4 xr[NV], dr[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * 32 + lane) * 8;
      xr[v] = __ldg(reinterpret_cast<const uint4*>(xg + row * p.ld_x + col));
      dr[v] = __ldg(reinterpret_cast<const uint4*>(dyg + row * p.ld_dy + col));
    }
    // two reduction rounds per row: the mean, then (sum of squares, sum g, sum g (x - mean)) together
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float x[8];
      unpack(xr[v], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[j];
    }
    const float mean = warp_sum(s) * kInvH;
    float q = 0.f, c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float x[8], dy[8], gam[8];
      unpack(xr[v], x);
      unpack(dr[v], dy);
      load8_f32(p.gamma + (v * 32 + lane) * 8, gam);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[j] - mean;
        const float g = dy[j] * gam[j];
        q = fmaf(d, d, q);
        c1 += g;
        c2 = fmaf(g, d, c2);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
      c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, o);
      c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
    }
    const float rstd = rsqrtf(q * kInvH + kLnEps);
    // dx = rstd (g - mean(g) - xh mean(g xh)) with xh = (x - mean) rstd, as  k1 g + k3 xh + k2
    const float k1 = rstd, k2 = -rstd * c1 * kInvH, k3 = -rstd * rstd * c2 * kInvH, m0 = -mean * rstd;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int col = (v * 32 + lane) * 8;
      float x[8], dy[8], gam[8], dx[8];
      unpack(xr[v], x);
      unpack(dr[v], dy);
      load8_f32(p.gamma + col, gam);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(x[j], rstd, m0);
        dx[j] = fmaf(k1, dy[j] * gam[j], fmaf(k3, xh, k2));
        ag[v][j] = fmaf(dy[j], xh, ag[v][j]);
        ab[v][j] += dy[j];
      }
      st8(p.dx, row * p.ld_dx + col, p.dx_f32, FMT, dx);
      if (masked) {
        const unsigned long long idx = static_cast<unsigned long long>(row) * H + col;
        const uint32_t inner = drop_inner(drop, idx);   // (H % 8 == 0: the 8 indices share their high word)
#pragma unroll
        for (int j = 0; j < 8; ++j) dx[j] = drop_keep(drop, inner, static_cast<uint32_t>(idx) + j) ? dx[j] * drop.inv_keep : 0.f;
        store8_16<FMT>(static_cast<uint16_t*>(p.dx_masked) + row * p.ld_dx + col, dx);
      }
      float4 a0 = *reinterpret_cast<float4*>(axs + col), a1 = *reinterpret_cast<float4*>(axs + col + 4);
      a0.x += dx[0]; a0.y += dx[1]; a0.z += dx[2]; a0.w += dx[3];
      a1.x += dx[4]; a1.y += dx[5]; a1.z += dx[6]; a1.w += dx[7];
      *reinterpret_cast<float4*>(axs + col) = a0;
      *reinterpret_cast<float4*>(axs + col + 4) = a1;
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    *reinterpret_cast<float4*>(mine + col) = make_float4(ag[v][0], ag[v][1], ag[v][2], ag[v][3]);
    *reinterpret_cast<float4*>(mine + col + 4) = make_float4(ag[v][4], ag[v][5], ag[v][6], ag[v][7]);
    *reinterpret_cast<float4*>(mine + H + col) = make_float4(ab[v][0], ab[v][1], ab[v][2], ab[v][3]);
    *reinterpret_cast<float4*>(mine + H + col + 4) = make_float4(ab[v][4], ab[v][5], ab[v][6], ab[v][7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ln_red[w * 3 * H + i];
    if (i < H) atomicAdd(p.dgamma + i, t);
    else if (i < 2 * H) atomicAdd(p.dbeta + i - H, t);
    else if (p.dxsum) atomicAdd(p.dxsum + i - 2 * H, t);
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

static unsigned row_grid(long long rows, int blocks_per_sm) {
  long long b = (rows + 7) / 8;
  // resident blocks only (72 KB of partial tables each at H = 768); every block ends with 3 H global atomics
  const long long cap = 148 * blocks_per_sm;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

int ln_bwd_run(const LnBwdParams& p, int H, void* stream) {
  LDOT_REQUIRE(p.rows >= 1 && H % 256 == 0, "layernorm_bwd: bad shape rows=%lld H=%d", p.rows, H);
  LDOT_REQUIRE(p.ld_dy % 8 == 0 && p.ld_x % 8 == 0 && p.ld_dx % 8 == 0, "layernorm_bwd: row pitches must be multiples of 8");
  const bool all16 = !p.x_f32 && !p.dy_f32;
  LDOT_REQUIRE(p.dx_masked == nullptr || (all16 && !p.dx_f32), "layernorm_bwd: the dropout form takes 16-bit dy, x and dx");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcLayerNorm, st, 0.0, static_cast<double>(p.rows) * H * (p.dx_masked ? 8.0 : 6.0));
  const size_t smem = static_cast<size_t>(8) * 3 * H * sizeof(float);
  LDOT_REQUIRE(smem <= 200 * 1024, "layernorm_bwd: hidden size %d too large", H);
#define LDOT_LNB_K(KERN, BPS) \
  { static bool cfgd = false; \
    if (!cfgd) { LDOT_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); cfgd = true; } \
    KERN<<<row_grid(p.rows, BPS), 256, smem, st>>>(p); }
  if (all16 && H <= 1024) {
    if (p.fmt == 1) { LDOT_NV_DISPATCH(H, LDOT_LNB_K((ln_bwd16_kernel<NV, 1>), 2)) }
    else { LDOT_NV_DISPATCH(H, LDOT_LNB_K((ln_bwd16_kernel<NV, 0>), 2)) }
  } else {
    LDOT_NV_DISPATCH(H, LDOT_LNB_K(ln_bwd_kernel<NV>, 1))
  }
#undef LDOT_LNB_K
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ attention backward

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
// grid (heads, B); block = SPAD / 16 warps.  qkv [B * S, 3 H] (saved forward activations), dctx [B * S, H],
// dqkv [B * S, 3 H].  All five contractions run on mma.sync m16n8k16 tiles (16-bit operands, fp32 accumulation):
//   phase 1   warp w owns QUERY rows [16 w, 16 w + 16): S = Q K^T / 8 + mask -> P = softmax(S) (recomputed exactly as the
//             forward does, never stored by it); dP = dO V^T; D_i = sum_j P_ij dP_ij (= dO_i . O_i); dS = P (dP - D) / 8;
//             dQ = dS K.  P and dS leave as 16-bit tiles in shared memory.
//   phase 2   warp w owns KEY rows [16 w, 16 w + 16): dV = P^T dO and dK = dS^T Q, the transposed A operands read with
//             ldmatrix.trans straight from the phase-1 tiles.
// Outputs are staged per warp and written with 16-byte coalesced stores.
template <int SPAD, int FMT>
__global__ void __launch_bounds__(SPAD * 2) attention_bwd_kernel(const uint16_t* __restrict__ qkv, const long long* __restrict__ mask,
                                                                  const uint16_t* __restrict__ dctx, uint16_t* __restrict__ dqkv,
                                                                  int S, int H, const DropKey drop_in) {
  const DropKey drop = drop_resolve(drop_in);
  constexpr int PP = SPAD + 8;          // pitch of the P / dS tiles (odd multiple of 16 B: conflict-free ldmatrix)
  constexpr int NT = SPAD / 8;
  extern __shared__ __align__(16) uint16_t ab_smem[];
  uint16_t* sQ = ab_smem;
  uint16_t* sK = sQ + SPAD * kRowPad;
  uint16_t* sV = sK + SPAD * kRowPad;
  uint16_t* sdO = sV + SPAD * kRowPad;
  uint16_t* sP = sdO + SPAD * kRowPad;
  uint16_t* sdS = sP + SPAD * PP;
  uint16_t* sOut = sdS + SPAD * PP;     // [warps][16][kRowPad] output staging
  float* sAdd = reinterpret_cast<float*>(sOut + (SPAD / 16) * 16 * kRowPad);   // additive mask per key column
  const int head = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tok0 = static_cast<long long>(b) * S;
  const int ld = 3 * H;

  for (int i = threadIdx.x; i < SPAD * 8 * 4; i += blockDim.x) {
    const int mat = i / (SPAD * 8), rem = i - mat * SPAD * 8;
    const int r = rem >> 3, c = (rem & 7) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < S) {
      const uint16_t* src = mat < 3 ? qkv + (tok0 + r) * ld + mat * H + head * kHeadDim + c
                                    : dctx + (tok0 + r) * H + head * kHeadDim + c;
      v = __ldg(reinterpret_cast<const uint4*>(src));
    }
    *reinterpret_cast<uint4*>(ab_smem + mat * SPAD * kRowPad + r * kRowPad + c) = v;
  }
  for (int j = threadIdx.x; j < SPAD; j += blockDim.x)
    sAdd[j] = j >= S ? -INFINITY : (mask[tok0 + j] != 0 ? 0.f : -10000.f);
  __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  const int row0 = warp * 16;
  uint16_t* stage = sOut + warp * 16 * kRowPad;
  auto smem_addr = [](const uint16_t* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); };
  // 16 x 64 fp32 fragments -> staging -> global rows [row0, row0 + 16) of column block `which` (0 = dQ, 1 = dK, 2 = dV)
  auto store_tile = [&](float (&o)[8][4], int which) {
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t*>(stage + g * kRowPad + n * 8 + t * 2) = pk2<FMT>(o[n][0], o[n][1]);
      *reinterpret_cast<uint32_t*>(stage + (g + 8) * kRowPad + n * 8 + t * 2) = pk2<FMT>(o[n][2], o[n][3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int r = idx >> 3, c = (idx & 7) * 8;
      if (row0 + r < S)
        *reinterpret_cast<uint4*>(dqkv + (tok0 + row0 + r) * ld + which * H + head * kHeadDim + c) =
            *reinterpret_cast<const uint4*>(stage + r * kRowPad + c);
    }
  };

  // ---------------------------------------------------------------- phase 1: query-row block of this warp
  {
    uint32_t qa[4][4], da[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int r = row0 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int c = ks * 16 + 8 * (lane >> 4);
      ldsm_x4(qa[ks], smem_addr(sQ + r * kRowPad + c));
      ldsm_x4(da[ks], smem_addr(sdO + r * kRowPad + c));
    }
    float sc[NT][4], dp[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
      dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
    }
#pragma unroll
    for (int n2 = 0; n2 < NT / 2; ++n2) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kb[4], vb[4];
        const int r = n2 * 16 + (lane & 7) + 8 * (lane >> 4);
        const int c = ks * 16 + 8 * ((lane >> 3) & 1);
        ldsm_x4(kb, smem_addr(sK + r * kRowPad + c));
        ldsm_x4(vb, smem_addr(sV + r * kRowPad + c));
        mma16816<FMT>(sc[2 * n2], qa[ks], kb[0], kb[1]);
        mma16816<FMT>(sc[2 * n2 + 1], qa[ks], kb[2], kb[3]);
        mma16816<FMT>(dp[2 * n2], da[ks], vb[0], vb[1]);
        mma16816<FMT>(dp[2 * n2 + 1], da[ks], vb[2], vb[3]);
      }
    }
    // softmax over keys, exactly as attention_kernel (rows g and g + 8 of this warp)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float add = sAdd[n * 8 + t * 2 + j];
        sc[n][j] = fmaf(sc[n][j], 0.125f, add);
        sc[n][2 + j] = fmaf(sc[n][2 + j], 0.125f, add);
        mx0 = fmaxf(mx0, sc[n][j]);
        mx1 = fmaxf(mx1, sc[n][2 + j]);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sc[n][j] = __expf(sc[n][j] - mx0);
        sc[n][2 + j] = __expf(sc[n][2 + j] - mx1);
        sum0 += sc[n][j];
        sum1 += sc[n][2 + j];
      }
    }
    sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, 1);
    sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, 2);
    sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, 1);
    sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    float d0 = 0.f, d1 = 0.f;
    // with attention dropout (mask M, keep scale c): O = (P o M c) V, so dP = (dO V^T) o M c.  mc() regenerates M c of
    // one element (hashing twice is cheaper than 4 NT live registers)
    const unsigned long long dbase = (static_cast<unsigned long long>(b) * gridDim.x + head) * S;
    auto mc = [&](int n, int e) -> float {   // e: 0, 1 = row g; 2, 3 = row g + 8
      const int col = n * 8 + t * 2 + (e & 1);
      return drop_keep(drop, (dbase + row0 + g + 8 * (e >> 1)) * S + col) ? drop.inv_keep : 0.f;
    };
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sc[n][j] *= inv0;
        sc[n][2 + j] *= inv1;
        if (drop.thr != 0) {
          dp[n][j] *= mc(n, j);
          dp[n][2 + j] *= mc(n, 2 + j);
        }
        d0 = fmaf(sc[n][j], dp[n][j], d0);
        d1 = fmaf(sc[n][2 + j], dp[n][2 + j], d1);
      }
    }
    d0 += __shfl_xor_sync(0xFFFFFFFFu, d0, 1);
    d0 += __shfl_xor_sync(0xFFFFFFFFu, d0, 2);
    d1 += __shfl_xor_sync(0xFFFFFFFFu, d1, 1);
    d1 += __shfl_xor_sync(0xFFFFFFFFu, d1, 2);
    // P and dS = P (dP - D) / 8 as 16-bit tiles: A fragments of dQ = dS K here, transposed operands of phase 2
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      // (dV = (P o M c)^T dO reads the DROPPED probabilities)
      uint32_t p_lo, p_hi;
      if (drop.thr != 0) {
        p_lo = pk2<FMT>(sc[n][0] * mc(n, 0), sc[n][1] * mc(n, 1));
        p_hi = pk2<FMT>(sc[n][2] * mc(n, 2), sc[n][3] * mc(n, 3));
      } else {
        p_lo = pk2<FMT>(sc[n][0], sc[n][1]);
        p_hi = pk2<FMT>(sc[n][2], sc[n][3]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        dp[n][j] = sc[n][j] * (dp[n][j] - d0) * 0.125f;
        dp[n][2 + j] = sc[n][2 + j] * (dp[n][2 + j] - d1) * 0.125f;
      }
      const int col = n * 8 + t * 2;
      *reinterpret_cast<uint32_t*>(sP + (row0 + g) * PP + col) = p_lo;
      *reinterpret_cast<uint32_t*>(sP + (row0 + g + 8) * PP + col) = p_hi;
      *reinterpret_cast<uint32_t*>(sdS + (row0 + g) * PP + col) = pk2<FMT>(dp[n][0], dp[n][1]);
      *reinterpret_cast<uint32_t*>(sdS + (row0 + g + 8) * PP + col) = pk2<FMT>(dp[n][2], dp[n][3]);
    }
#pragma unroll
    for (int kk = 0; kk < SPAD / 16; ++kk) {
      uint32_t sa[4];
      sa[0] = pk2<FMT>(dp[2 * kk][0], dp[2 * kk][1]);
      sa[1] = pk2<FMT>(dp[2 * kk][2], dp[2 * kk][3]);
      sa[2] = pk2<FMT>(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      sa[3] = pk2<FMT>(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t kb[4];
        const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int c = d2 * 16 + 8 * (lane >> 4);
        ldsm_x4_t(kb, smem_addr(sK + r * kRowPad + c));
        mma16816<FMT>(o[2 * d2], sa, kb[0], kb[1]);
        mma16816<FMT>(o[2 * d2 + 1], sa, kb[2], kb[3]);
      }
    }
    store_tile(o, 0);
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase 2: key-row block of this warp
  {
    float ov[8][4], ok[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      ov[n][0] = ov[n][1] = ov[n][2] = ov[n][3] = 0.f;
      ok[n][0] = ok[n][1] = ok[n][2] = ok[n][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < SPAD / 16; ++kk) {
      // A = (P^T | dS^T)[keys row0 .. +16, queries 16 kk .. +16]: tile (m, k) sits at stored [16 kk + k][row0 + m]
      uint32_t pa[4], sa[4];
      const int ar = kk * 16 + (lane & 7) + 8 * (lane >> 4);
      const int ac = row0 + 8 * ((lane >> 3) & 1);
      ldsm_x4_t(pa, smem_addr(sP + ar * PP + ac));
      ldsm_x4_t(sa, smem_addr(sdS + ar * PP + ac));
#pragma unroll
      for (int d2 = 0; d2 < 4; ++d2) {
        uint32_t ob[4], qb[4];
        const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int c = d2 * 16 + 8 * (lane >> 4);
        ldsm_x4_t(ob, smem_addr(sdO + r * kRowPad + c));
        ldsm_x4_t(qb, smem_addr(sQ + r * kRowPad + c));
        mma16816<FMT>(ov[2 * d2], pa, ob[0], ob[1]);
        mma16816<FMT>(ov[2 * d2 + 1], pa, ob[2], ob[3]);
        mma16816<FMT>(ok[2 * d2], sa, qb[0], qb[1]);
        mma16816<FMT>(ok[2 * d2 + 1], sa, qb[2], qb[3]);
      }
    }
    store_tile(ok, 1);
    store_tile(ov, 2);
  }
}

template <int FMT>
static int attention_bwd_launch(const void* qkv, const long long* mask, const void* dctx, void* dqkv, int B, int S, int H,
                                int heads, const DropKey& drop, cudaStream_t st) {
  const int spad = (S + 15) / 16 * 16;
  const dim3 grid(heads, B);
  const size_t smem = (static_cast<size_t>(4) * spad * kRowPad + static_cast<size_t>(2) * spad * (spad + 8) +
                       static_cast<size_t>(spad / 16) * 16 * kRowPad) * sizeof(uint16_t) + static_cast<size_t>(spad) * sizeof(float);
#define LDOT_ATTB_CASE(SP)                                                                                         \
  case SP: {                                                                                                       \
    auto kern = attention_bwd_kernel<SP, FMT>;                                                                     \
    if (smem > 48 * 1024) LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, SP * 2, smem, st>>>(static_cast<const uint16_t*>(qkv), mask, static_cast<const uint16_t*>(dctx),  \
                                      static_cast<uint16_t*>(dqkv), S, H, drop);                                   \
    break;                                                                                                         \
  }
  switch (spad) {
    LDOT_ATTB_CASE(16)
    LDOT_ATTB_CASE(32)
    LDOT_ATTB_CASE(48)
    LDOT_ATTB_CASE(64)
    LDOT_ATTB_CASE(80)
    LDOT_ATTB_CASE(96)
    LDOT_ATTB_CASE(112)
    LDOT_ATTB_CASE(128)
    default:
      return set_error(kErrArg, "attention_bwd: sequence length %d > 128 not supported", S);
  }
#undef LDOT_ATTB_CASE
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int attention_bwd_run(const void* qkv, const long long* mask, const void* ctx, const void* dctx, void* dqkv, int B, int S,
                      int H, int heads, int fmt, void* stream, float drop_p, unsigned long long seed, int site) {
  (void)ctx;   // (D_i = dO_i . O_i is evaluated as sum_j P_ij dP_ij from the recomputed probabilities)
  LDOT_REQUIRE(B >= 1 && S >= 1 && S <= 128, "attention_bwd: bad shape B=%d S=%d (S <= 128)", B, S);
  LDOT_REQUIRE(H == heads * kHeadDim, "attention_bwd: hidden %d must be heads (%d) x 64", H, heads);
  LDOT_REQUIRE(B <= 65535, "attention_bwd: batch %d > 65535 (split the batch)", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcAttention, st, 10.0 * B * static_cast<double>(S) * S * H, static_cast<double>(B) * S * H * 16.0);
  LDOT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attention_bwd: dropout probability %f out of [0, 1)", drop_p);
  const DropKey drop = make_drop_key(drop_p, seed, site);
  return fmt == 1 ? attention_bwd_launch<1>(qkv, mask, dctx, dqkv, B, S, H, heads, drop, st)
                  : attention_bwd_launch<0>(qkv, mask, dctx, dqkv, B, S, H, heads, drop, st);
}

static unsigned flat_grid(long long n_vec) {
  long long b = (n_vec + 255) / 256;
  const long long cap = 148 * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------ dropout (elementwise)
// out = dropout(x) (+ res): hidden-state dropout of the towers (layer.py:113,154; model.py:245,272) applied to a 16-bit
// [rows, cols] matrix (row pitch ld, cols % 8 == 0); the same call masks the gradient in backward (res = null).
__global__ void __launch_bounds__(256) dropout_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ res,
                                                      uint16_t* __restrict__ out, long long rows, int cols, long long ld,
                                                      int fmt, const DropKey drop_in) {
  const DropKey drop = drop_resolve(drop_in);
  const int c8 = cols / 8;
  const long long n8 = rows * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c8;
    const int c = static_cast<int>(i - r * c8) * 8;
    float f[8];
    ld8(x, r * ld + c, 0, fmt, f);
    const unsigned long long idx = static_cast<unsigned long long>(r) * cols + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = drop_keep(drop, idx + j) ? f[j] * drop.inv_keep : 0.f;
    if (res != nullptr) {
      float g[8];
      ld8(res, r * ld + c, 0, fmt, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += g[j];
    }
    st8(out, r * ld + c, 0, fmt, f);
  }
}

int dropout_run(const void* x, const void* res, void* out, long long rows, int cols, long long ld, float p,
                unsigned long long seed, int site, int fmt, void* stream) {
  LDOT_REQUIRE(rows >= 0 && cols >= 8 && cols % 8 == 0 && ld % 8 == 0 && ld >= cols, "dropout: bad shape rows=%lld cols=%d", rows, cols);
  LDOT_REQUIRE(p > 0.f && p < 1.f, "dropout: probability %f out of (0, 1)", p);
  if (rows == 0) return kOk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcCast, st, 0.0, static_cast<double>(rows) * cols * (res ? 6.0 : 4.0));
  dropout_kernel<<<flat_grid(rows * (cols / 8)), 256, 0, st>>>(static_cast<const uint16_t*>(x), static_cast<const uint16_t*>(res),
                                                             static_cast<uint16_t*>(out), rows, cols, ld, fmt,
                                                             make_drop_key(p, seed, site));
  LDOT_CHECK_LAUNCH();
  return kOk;
}

// ------------------------------------------------------------------------------------------------ GELU (elementwise)
// mode 0: out = gelu(x);  mode 1: out = dy * gelu'(x);  mode 2 (training forward): out = gelu(x) and x <- gelu'(x) IN PLACE
// (dy = x's own buffer): backward needs nothing else of the pre-activation, and its dgrad epilogue becomes one multiply.
// Bandwidth-bound here (2 B read + 4 B written per element) - the same math inside the K = 768 GEMM epilogues made them
// issue-bound: FFN-up forward 151 us fused vs 60 + 51 us as GEMM + this kernel (17.6 k tokens).
__global__ void __launch_bounds__(256) gelu_kernel(const uint16_t* __restrict__ x, const uint16_t* dy,
                                                   uint16_t* __restrict__ out, long long n8, int mode, int fmt) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8], g[8];
    ld8(x, i * 8, 0, fmt, f);
    if (mode == 1) {
      ld8(dy, i * 8, 0, fmt, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = g[j] * gelu_erf_grad(f[j]);
    } else if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) gelu_erf_both(f[j], f[j], g[j]);
      st8(const_cast<uint16_t*>(dy), i * 8, 0, fmt, g);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
    }
    st8(out, i * 8, 0, fmt, f);
  }
}


int gelu_run(const void* x, const void* dy, void* out, long long n, int mode, int fmt, void* stream) {
  LDOT_REQUIRE(n >= 0 && n % 8 == 0, "gelu: element count %lld must be a multiple of 8", n);
  LDOT_REQUIRE((mode >= 1) == (dy != nullptr), "gelu: the second buffer is required by (and only by) modes 1 and 2");
  if (n == 0) return kOk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcCast, st, 0.0, static_cast<double>(n) * (mode ? 6.0 : 4.0));
  // (mode 2: x and dy are the same buffer - read, then overwritten by the same thread)
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
  float lr = a.lr, bc1 = a.bc1, bc2_sqrt = a.bc2_sqrt;
  if (a.hyper != nullptr) {
    lr = __ldg(a.hyper);
    bc1 = __ldg(a.hyper + 1);
    bc2_sqrt = __ldg(a.hyper + 2);
  }
  const float step = lr / bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float g = a.g[i] * clip;
    float p = a.p[i] * (1.f - lr * a.wd);
    const float m = a.beta1 * a.m[i] + (1.f - a.beta1) * g;
    const float v = a.beta2 * a.v[i] + (1.f - a.beta2) * g * g;
    p -= step * m / (sqrtf(v) / bc2_sqrt + a.eps);
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
