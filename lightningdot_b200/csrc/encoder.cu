// Non-GEMM pieces of the two towers (BERT-base text tower, UNITER-base image-region tower):
//   embed_text     LN(word[ids] + pos[position_ids] + type[0])                 uniter_model/model/model.py:233-246
//   embed_image    LN(LN(img_linear(feat)) + LN(pos_linear(box)) + type[1])    uniter_model/model/model.py:262-273,328-336
//   layernorm      row LayerNorm, eps 1e-12 (after the residual GEMM epilogue)  uniter_model/model/layer.py:111-115,152-156
//   attention      softmax(Q K^T / 8 + (1 - mask) * -10000) V, 12 heads x 64    uniter_model/model/layer.py:80-101
//   cast           fp32 -> 16-bit (region features before img_linear)
// Activations are 16-bit (bf16 default, fp16 selectable: `fmt` 1 / 0), statistics and softmax in fp32.
// Sequences are tiny (text <= 62, image <= 101 positions), so attention is one CTA per (sequence, head) on
// mma.sync tiles: < 1 % of the tower FLOPs, bandwidth-bound; the dense contractions (> 97 % of FLOPs) are the
// tcgen05 GEMMs in gemm_ops.cu.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include "host_common.h"
#include "prof.h"
#include "encoder_params.h"
#include "rowops.cuh"
#include "mma_sync.cuh"
#include "dropout.cuh"

namespace ldot {

// In-register LayerNorm of one row held as NV vectors of 8 per lane (vector v covers columns (v*32 + lane)*8 ..+7).
template <int NV>
__device__ __forceinline__ void warp_layernorm(float (&x)[NV][8], int H, int lane, const float* __restrict__ gamma,
                                               const float* __restrict__ beta) {
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

// ------------------------------------------------------------------------------------------------ layernorm
// one warp per row; H = NV * 256
template <int NV, int FMT, bool IN_F32>
__global__ void __launch_bounds__(256) layernorm_kernel(const void* __restrict__ in, long long ld_in,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        uint16_t* __restrict__ out, long long ld_out, long long rows) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float x[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    if (IN_F32) load8_f32(static_cast<const float*>(in) + row * ld_in + col, x[v]);
    else load8_16<FMT>(static_cast<const uint16_t*>(in) + row * ld_in + col, x[v]);
  }
  warp_layernorm<NV>(x, NV * 256, lane, gamma, beta);
#pragma unroll
  for (int v = 0; v < NV; ++v) store8_16<FMT>(out + row * ld_out + (v * 32 + lane) * 8, x[v]);
}

// ------------------------------------------------------------------------------------------------ text embeddings
// token (b, l) -> out row b * out_seq + l;  tables are 16-bit [rows, H]
template <int NV, int FMT>
__global__ void __launch_bounds__(256) embed_text_kernel(const long long* __restrict__ ids, const long long* __restrict__ pos_ids,
                                                         long long pos_batch_stride, const uint16_t* __restrict__ word,
                                                         const uint16_t* __restrict__ pos, const uint16_t* __restrict__ type0,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         uint16_t* __restrict__ out, int B, int L, int out_seq, int vocab,
                                                         int max_pos) {
  constexpr int H = NV * 256;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(B) * L) return;
  const int b = static_cast<int>(tok / L), l = static_cast<int>(tok - static_cast<long long>(b) * L);
  long long id = ids[tok];
  long long pid = pos_ids[b * pos_batch_stride + l];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  pid = pid < 0 ? 0 : (pid >= max_pos ? max_pos - 1 : pid);
  float x[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    float w[8], p[8], t[8];
    load8_16<FMT>(word + id * H + col, w);
    load8_16<FMT>(pos + pid * H + col, p);
    load8_16<FMT>(type0 + col, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[v][j] = w[j] + p[j] + t[j];
  }
  warp_layernorm<NV>(x, H, lane, gamma, beta);
  uint16_t* o = out + (static_cast<long long>(b) * out_seq + l) * H;
#pragma unroll
  for (int v = 0; v < NV; ++v) store8_16<FMT>(o + (v * 32 + lane) * 8, x[v]);
}

// ------------------------------------------------------------------------------------------------ image embeddings

template <int NV, int FMT>
__global__ void __launch_bounds__(256) embed_image_kernel(const EmbedImageParams p) {
  constexpr int H = NV * 256;
  const long long tok = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= static_cast<long long>(p.B) * p.R) return;
  const int b = static_cast<int>(tok / p.R), r = static_cast<int>(tok - static_cast<long long>(b) * p.R);
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
  }
  warp_layernorm<NV>(a, H, lane, p.img_g, p.img_b);
  warp_layernorm<NV>(q, H, lane, p.pos_g, p.pos_b);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int col = (v * 32 + lane) * 8;
    float t[8];
    load8_f32(p.type1 + col, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[v][j] = a[v][j] + q[v][j] + t[j];
  }
  warp_layernorm<NV>(a, H, lane, p.ln_g, p.ln_b);
  uint16_t* o = p.out + (static_cast<long long>(b) * p.out_seq + p.row_offset + r) * H;
#pragma unroll
  for (int v = 0; v < NV; ++v) store8_16<FMT>(o + (v * 32 + lane) * 8, a[v]);
}

// ------------------------------------------------------------------------------------------------ cast
template <int FMT>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, long long n8) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    load8_f32(in + i * 8, f);
    store8_16<FMT>(out + i * 8, f);
  }
}

// ------------------------------------------------------------------------------------------------ attention
// One (sequence, head) item: block = (SPAD / 16) warps; warp w owns query rows [16 w, 16 w + 16).  Only the first q_rows query
// positions of each sequence are computed and written (ctx is [B * q_rows, H]): q_rows = S for a full layer, 1 for the
// last layer of a tower whose caller only reads the [CLS] row (dvl/models/bi_encoder.py:120,188).
template <int SPAD, int FMT>
__device__ __forceinline__ void attention_item(uint16_t* sQ, const uint16_t* sK, const uint16_t* sV,
                                               const long long* smask, uint16_t* __restrict__ ctx, int b,
                                               int head, int heads, long long tok0, int S, int H, int q_rows,
                                               const DropKey& drop, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int qrow0 = warp * 16;
  if (qrow0 >= q_rows) return;  // (the caller synchronises the block after every item)
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int r = qrow0 + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int c = ks * 16 + 8 * (lane >> 4);
    ldsm_x4(qa[ks], static_cast<uint32_t>(__cvta_generic_to_shared(sQ + r * kRowPad + c)));
  }
  constexpr int NT = SPAD / 8;
  float sc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
  for (int n2 = 0; n2 < NT / 2; ++n2) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t kb[4];
      const int r = n2 * 16 + (lane & 7) + 8 * (lane >> 4);
      const int c = ks * 16 + 8 * ((lane >> 3) & 1);
      ldsm_x4(kb, static_cast<uint32_t>(__cvta_generic_to_shared(sK + r * kRowPad + c)));
      mma16816<FMT>(sc[2 * n2], qa[ks], kb[0], kb[1]);
      mma16816<FMT>(sc[2 * n2 + 1], qa[ks], kb[2], kb[3]);
    }
  }
  // scale, additive mask (uniter_model/model/model.py:362-365), softmax over keys (rows g and g + 8 of this warp)
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = n * 8 + t * 2 + j;
      float add;
      if (col >= S) add = -INFINITY;
      else add = smask[col] != 0 ? 0.f : -10000.f;   // (the sequence's attention_mask row, staged with Q / K / V)
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

  if (drop.thr != 0) {
    // attention-probability dropout (training, uniter_model/model/layer.py:93): dropped probabilities leave the P V
    // product, the survivors' 1 / (1 - p) is folded into the final normalisation
    const unsigned long long base = (static_cast<unsigned long long>(b) * heads + head) * S;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = n * 8 + t * 2 + j;
        if (!drop_keep(drop, (base + qrow0 + g) * S + col)) sc[n][j] = 0.f;
        if (!drop_keep(drop, (base + qrow0 + g + 8) * S + col)) sc[n][2 + j] = 0.f;
      }
    }
  }
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < SPAD / 16; ++kk) {
    uint32_t pa[4];
    pa[0] = pk2<FMT>(sc[2 * kk][0], sc[2 * kk][1]);
    pa[1] = pk2<FMT>(sc[2 * kk][2], sc[2 * kk][3]);
    pa[2] = pk2<FMT>(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
    pa[3] = pk2<FMT>(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
    for (int d2 = 0; d2 < 4; ++d2) {
      uint32_t vb[4];
      const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int c = d2 * 16 + 8 * (lane >> 4);
      ldsm_x4_t(vb, static_cast<uint32_t>(__cvta_generic_to_shared(sV + r * kRowPad + c)));
      mma16816<FMT>(o[2 * d2], pa, vb[0], vb[1]);
      mma16816<FMT>(o[2 * d2 + 1], pa, vb[2], vb[3]);
    }
  }
  const float inv0 = drop.inv_keep / sum0, inv1 = drop.inv_keep / sum1;
  // stage this warp's 16 x 64 output in its own (now dead) Q rows, then 16-byte coalesced stores
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(sQ + (qrow0 + g) * kRowPad + n * 8 + t * 2) = pk2<FMT>(o[n][0] * inv0, o[n][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + (qrow0 + g + 8) * kRowPad + n * 8 + t * 2) = pk2<FMT>(o[n][2] * inv1, o[n][3] * inv1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    const int r = qrow0 + (idx >> 3), c = (idx & 7) * 8;
    if (r < q_rows)
      *reinterpret_cast<uint4*>(ctx + (static_cast<long long>(b) * q_rows + r) * H + head * kHeadDim + c) =
          *reinterpret_cast<const uint4*>(sQ + r * kRowPad + c);
  }
}


__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}

// PERSISTENT over (sequence, head) items: a grid of resident CTAs strides over the B * heads items with two
// shared-memory buffers - the Q / K / V head slices of the next item arrive by cp.async while the current item is
// computed.  (One CTA per item made 120 000 64-thread CTAs per launch at 10 000 captions: launch-rate bound.)
template <int SPAD, int FMT>
__global__ void __launch_bounds__(SPAD * 2) attention_kernel(const uint16_t* __restrict__ qkv, const long long* __restrict__ mask,
                                                              uint16_t* __restrict__ ctx, int B, int S, int H, int heads,
                                                              int q_rows, const DropKey drop_in, int bufs) {
  const DropKey drop = drop_resolve(drop_in);
  extern __shared__ __align__(16) uint16_t att_smem_all[];
  constexpr int kBuf = 3 * SPAD * kRowPad + SPAD * 4;   // 16-bit elements per buffer: Q | K | V | int64 mask row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 3 * H;
  const int total = B * heads;

  // stage Q, K, V head slices [S, 64] of `item` (rows >= S zero-filled): 8 lanes x 16 B per row
  auto prefetch = [&](int item, int buf) {
    const int b_ = item / heads, head_ = item - b_ * heads;
    const long long tok0_ = static_cast<long long>(b_) * S;
    uint16_t* dst = att_smem_all + buf * kBuf;
    for (int i = threadIdx.x; i < SPAD * 8 * 3; i += blockDim.x) {
      const int mat = i / (SPAD * 8), rem = i - mat * SPAD * 8;
      const int r = rem >> 3, c = (rem & 7) * 8;
      const int rs = r < S ? r : S - 1;
      cp_async16_zfill(dst + mat * SPAD * kRowPad + r * kRowPad + c,
                       qkv + (tok0_ + rs) * ld + mat * H + head_ * kHeadDim + c, r < S);
    }
    if (threadIdx.x < SPAD) {   // attention_mask[b, :] (int64): read right after the score MMAs, so it rides along
      const int j = threadIdx.x < S ? threadIdx.x : S - 1;
      const uint32_t dstm = static_cast<uint32_t>(__cvta_generic_to_shared(dst + 3 * SPAD * kRowPad + threadIdx.x * 4));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dstm), "l"(mask + tok0_ + j) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // bufs == 2: the next item's slices arrive while this one is computed; bufs == 1: one buffer, twice the resident CTAs
  // per SM (the loads of other CTAs overlap instead)
  int buf = 0;
  if (bufs == 2 && static_cast<int>(blockIdx.x) < total) prefetch(blockIdx.x, 0);
  for (int item = blockIdx.x; item < total; item += gridDim.x, buf ^= (bufs - 1)) {
    const int nxt = item + gridDim.x;
    if (bufs == 1) {
      prefetch(item, 0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if (nxt < total) {
      prefetch(nxt, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int b = item / heads, head = item - b * heads;
    const long long tok0 = static_cast<long long>(b) * S;
    uint16_t* sQ = att_smem_all + buf * kBuf;
    uint16_t* sK = sQ + SPAD * kRowPad;
    uint16_t* sV = sK + SPAD * kRowPad;
    attention_item<SPAD, FMT>(sQ, sK, sV, reinterpret_cast<const long long*>(sV + SPAD * kRowPad), ctx, b, head, heads,
                              tok0, S, H, q_rows, drop, warp, lane);
    __syncthreads();   // every warp is done with this buffer before the prefetch of item + 2 * gridDim.x lands in it
  }
}


// ================================================================================================ host side
template <int FMT>
static int attention_launch(const void* qkv, const long long* mask, void* ctx, int B, int S, int H, int heads, int q_rows,
                            const DropKey& drop, cudaStream_t st) {
  const int spad = (S + 15) / 16 * 16;
  static int bufs = 0;
  if (bufs == 0) {
    const char* e = getenv("LDOT_ATT_BUFS");   // (measurement switch; one buffer measured ~4 % faster per clock at L = 32)
    bufs = e && atoi(e) == 2 ? 2 : 1;
  }
  const size_t smem = static_cast<size_t>(bufs) * (3 * spad * kRowPad + spad * 4) * sizeof(uint16_t);
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  long long per_sm = (200 * 1024) / static_cast<long long>(smem);
  per_sm = per_sm < 1 ? 1 : (per_sm > 16 ? 16 : per_sm);
  const long long items = static_cast<long long>(B) * heads;
  const unsigned grid = static_cast<unsigned>(items < per_sm * sms ? items : per_sm * sms);
#define LDOT_ATT_CASE(SP)                                                                                         \
  case SP: {                                                                                                      \
    auto kern = attention_kernel<SP, FMT>;                                                                        \
    if (smem > 48 * 1024) LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, SP * 2, smem, st>>>(static_cast<const uint16_t*>(qkv), mask, static_cast<uint16_t*>(ctx), B, S, H, heads, q_rows, drop, bufs);  \
    break;                                                                                                        \
  }
  switch (spad) {
    LDOT_ATT_CASE(16)
    LDOT_ATT_CASE(32)
    LDOT_ATT_CASE(48)
    LDOT_ATT_CASE(64)
    LDOT_ATT_CASE(80)
    LDOT_ATT_CASE(96)
    LDOT_ATT_CASE(112)
    LDOT_ATT_CASE(128)
    default:
      return set_error(kErrArg, "attention: sequence length %d > 128 not supported", S);
  }
#undef LDOT_ATT_CASE
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int attention_run(const void* qkv, const long long* mask, void* ctx, int B, int S, int H, int heads, int q_rows, int fmt,
                  void* stream, float drop_p, unsigned long long seed, int site) {
  LDOT_REQUIRE(B >= 1 && S >= 1 && S <= 128, "attention: bad shape B=%d S=%d (S <= 128)", B, S);
  LDOT_REQUIRE(q_rows >= 1 && q_rows <= S, "attention: q_rows %d must be in [1, S = %d]", q_rows, S);
  LDOT_REQUIRE(H == heads * kHeadDim && H % 8 == 0, "attention: hidden %d must be heads (%d) x 64", H, heads);
  LDOT_REQUIRE(B <= 65535, "attention: batch %d > 65535 (split the batch)", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcAttention, st, 4.0 * B * static_cast<double>(q_rows) * S * H,
                 static_cast<double>(B) * (static_cast<double>(S) * H * 4.0 + static_cast<double>(q_rows) * H * 4.0));
  LDOT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attention: dropout probability %f out of [0, 1)", drop_p);
  const DropKey drop = make_drop_key(drop_p, seed, site);
  return fmt == 1 ? attention_launch<1>(qkv, mask, ctx, B, S, H, heads, q_rows, drop, st)
                  : attention_launch<0>(qkv, mask, ctx, B, S, H, heads, q_rows, drop, st);
}

#define LDOT_NV_DISPATCH(H, CALL)                                   \
  switch ((H) / 256) {                                              \
    case 1: { constexpr int NV = 1; CALL; break; }                  \
    case 2: { constexpr int NV = 2; CALL; break; }                  \
    case 3: { constexpr int NV = 3; CALL; break; }                  \
    case 4: { constexpr int NV = 4; CALL; break; }                  \
    case 6: { constexpr int NV = 6; CALL; break; }                  \
    case 8: { constexpr int NV = 8; CALL; break; }                  \
    default: return set_error(kErrArg, "hidden size %d not supported (256 x {1,2,3,4,6,8})", (H)); \
  }

int layernorm_run(const void* in, long long ld_in, int in_f32, const float* gamma, const float* beta, void* out,
                  long long ld_out, long long rows, int H, int fmt, void* stream) {
  LDOT_REQUIRE(rows >= 1 && H % 256 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0, "layernorm: bad shape rows=%lld H=%d", rows, H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  uint16_t* o = static_cast<uint16_t*>(out);
  KernelScope ks(kKcLayerNorm, st, 0.0, static_cast<double>(rows) * H * (in_f32 ? 6.0 : 4.0));
  if (fmt == 1) {
    if (in_f32) { LDOT_NV_DISPATCH(H, (layernorm_kernel<NV, 1, true><<<blocks, 256, 0, st>>>(in, ld_in, gamma, beta, o, ld_out, rows))) }
    else { LDOT_NV_DISPATCH(H, (layernorm_kernel<NV, 1, false><<<blocks, 256, 0, st>>>(in, ld_in, gamma, beta, o, ld_out, rows))) }
  } else {
    if (in_f32) { LDOT_NV_DISPATCH(H, (layernorm_kernel<NV, 0, true><<<blocks, 256, 0, st>>>(in, ld_in, gamma, beta, o, ld_out, rows))) }
    else { LDOT_NV_DISPATCH(H, (layernorm_kernel<NV, 0, false><<<blocks, 256, 0, st>>>(in, ld_in, gamma, beta, o, ld_out, rows))) }
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int embed_text_run(const long long* ids, const long long* pos_ids, long long pos_batch_stride, const void* word,
                   const void* pos, const void* type0, const float* gamma, const float* beta, void* out, int B, int L,
                   int out_seq, int H, int vocab, int max_pos, int fmt, void* stream) {
  LDOT_REQUIRE(B >= 1 && L >= 1 && out_seq >= L && H % 256 == 0, "embed_text: bad shape B=%d L=%d H=%d", B, L, H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(B) * L + 7) / 8);
  const uint16_t* w = static_cast<const uint16_t*>(word);
  const uint16_t* p = static_cast<const uint16_t*>(pos);
  const uint16_t* t = static_cast<const uint16_t*>(type0);
  uint16_t* o = static_cast<uint16_t*>(out);
  KernelScope ks(kKcEmbed, st, 0.0, static_cast<double>(B) * L * H * 6.0);
  if (fmt == 1) { LDOT_NV_DISPATCH(H, (embed_text_kernel<NV, 1><<<blocks, 256, 0, st>>>(ids, pos_ids, pos_batch_stride, w, p, t, gamma, beta, o, B, L, out_seq, vocab, max_pos))) }
  else { LDOT_NV_DISPATCH(H, (embed_text_kernel<NV, 0><<<blocks, 256, 0, st>>>(ids, pos_ids, pos_batch_stride, w, p, t, gamma, beta, o, B, L, out_seq, vocab, max_pos))) }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int embed_image_run(const EmbedImageParams& p, int H, int fmt, void* stream) {
  LDOT_REQUIRE(p.B >= 1 && p.R >= 1 && H % 256 == 0 && p.out_seq >= p.row_offset + p.R, "embed_image: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((static_cast<long long>(p.B) * p.R + 7) / 8);
  KernelScope ks(kKcEmbed, st, 0.0, static_cast<double>(p.B) * p.R * H * 6.0);
  if (fmt == 1) { LDOT_NV_DISPATCH(H, (embed_image_kernel<NV, 1><<<blocks, 256, 0, st>>>(p))) }
  else { LDOT_NV_DISPATCH(H, (embed_image_kernel<NV, 0><<<blocks, 256, 0, st>>>(p))) }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int cast_run(const float* in, void* out, long long n, int fmt, void* stream) {
  LDOT_REQUIRE(n >= 0 && n % 8 == 0, "cast: element count %lld must be a multiple of 8", n);
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "cast: alignment");
  if (n == 0) return kOk;
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcCast, st, 0.0, static_cast<double>(n) * 6.0);
  if (fmt == 1) cast_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, static_cast<uint16_t*>(out), n8);
  else cast_kernel<0><<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, static_cast<uint16_t*>(out), n8);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

}  // namespace ldot
