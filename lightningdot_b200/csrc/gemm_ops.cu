// Encoder Linear layers on the tcgen05 main loop:  out = act(A . W^T + bias) (+ residual)
//
// Replaces every nn.Linear on the tower path (uniter_model/model/layer.py:64-66,76-78,107,133,148;
// uniter_model/model/model.py:252; dvl/models/bi_encoder.py:83-88,138-143).  A = activations [M, K] (16-bit,
// K-major), W = nn.Linear weight [N, K] (16-bit, K-major, used as stored - no transpose), fp32 accumulate in TMEM,
// bias / erf-GELU / residual add fused in the epilogue (one thread = one output row), 16-bit or fp32 output.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "gemm_tc.cuh"
#include "host_common.h"
#include "prof.h"

namespace ldot {

constexpr int kLinBN = 256;
constexpr int kLinStages = 4;

struct StoreParams {
  void* out;               // [M, ldo] bf16/fp16 or fp32
  const float* bias;       // [N] or null
  const void* residual;    // [M, ldr] same 16-bit type as A, or null
  long long ldo, ldr;
  long long M;
  int N;
  int act;                 // 0 = identity, 1 = erf-GELU
  int out_f32;             // 1: fp32 output, 0: 16-bit output of type `fmt`
  int fmt;                 // 0 = fp16, 1 = bf16
};

// erf with |abs error| <= 1.5e-7 (Abramowitz & Stegun 7.1.26) - far below the 16-bit output resolution, and a
// third of the instructions of erff(), which matters because the FFN-up epilogue is ALU-paced.
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f)); }

__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
  if (fmt == 1) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, int fmt) {
  if (fmt == 1) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

struct EpiStore {
  using Params = StoreParams;
  struct State {};
  static __device__ __forceinline__ void unit_begin(State&, const Params&, const UnitInfo&, int) {}
  static __device__ __forceinline__ void unit_end(State&, const Params&, const UnitInfo&, int) {}

  static __device__ __forceinline__ void tile(State&, const Params& p, const UnitInfo& u, int row, int nt,
                                              uint32_t taddr) {
    const long long grow = static_cast<long long>(u.m_tile) * kBM + row;
    const bool row_ok = grow < p.M;
#pragma unroll 1
    for (int c = 0; c < kLinBN; c += 32) {
      const int col = nt * kLinBN + c;
      if (col >= p.N) break;  // uniform across the warp
      uint32_t v[32];
      ptx::tmem_ld32(taddr + c, v);
      ptx::tmem_ld_wait();
      if (!row_ok) continue;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const bool full = col + 32 <= p.N;
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || col + j < p.N) f[j] += __ldg(p.bias + col + j);
      }
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
      }
      if (p.residual) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(p.residual) + grow * p.ldr + col);
        if (full) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 rr = __ldg(reinterpret_cast<const uint4*>(r) + j4);
            const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 x = unpack2(w[t], p.fmt);
              f[j4 * 8 + t * 2] += x.x;
              f[j4 * 8 + t * 2 + 1] += x.y;
            }
          }
        } else {
          const uint16_t* r16 = static_cast<const uint16_t*>(p.residual) + grow * p.ldr + col;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col + j < p.N) f[j] += unpack2(static_cast<uint32_t>(r16[j]), p.fmt).x;
        }
      }
      if (p.out_f32) {
        float* o = static_cast<float*>(p.out) + grow * p.ldo + col;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col + j < p.N) o[j] = f[j];
        }
      } else {
        uint16_t* o = static_cast<uint16_t*>(p.out) + grow * p.ldo + col;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 w;
            w.x = pack2(f[j], f[j + 1], p.fmt);
            w.y = pack2(f[j + 2], f[j + 3], p.fmt);
            w.z = pack2(f[j + 4], f[j + 5], p.fmt);
            w.w = pack2(f[j + 6], f[j + 7], p.fmt);
            *reinterpret_cast<uint4*>(o + j) = w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col + j < p.N) o[j] = static_cast<uint16_t>(pack2(f[j], 0.f, p.fmt) & 0xFFFFu);
        }
      }
    }
  }
};

int linear_run(const void* a, long long lda, const void* w, long long ldw, const float* bias, const void* residual,
               long long ldr, void* out, long long ldo, long long M, int N, int K, int fmt, int act, int out_f32,
               void* stream) {
  LDOT_REQUIRE(M >= 1 && N >= 1 && K >= 8, "bad GEMM shape M=%lld N=%d K=%d", M, N, K);
  LDOT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "K, lda, ldw must be multiples of 8 elements");
  LDOT_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(out_f32 ? (ldo % 4 == 0) : (ldo % 8 == 0), "ldo alignment");
  LDOT_REQUIRE(!residual || ldr % 8 == 0, "ldr alignment");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
               "out / residual must be 16-byte aligned");
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  CUtensorMap ta, tb;
  if (int e = make_tmap_kmajor_16b(&ta, a, M, K, static_cast<uint64_t>(lda) * 2, kBM)) return e;
  if (int e = make_tmap_kmajor_16b(&tb, w, N, K, static_cast<uint64_t>(ldw) * 2, kLinBN)) return e;
  GemmSched s;
  s.m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  s.n_tiles = (N + kLinBN - 1) / kLinBN;
  s.tiles_per_unit = s.n_tiles;  // one unit = one 128-row block x all N tiles (A tile reused from L2)
  s.chunks = 1;
  s.num_units = s.m_tiles;
  s.k_blocks = (K + kBK - 1) / kBK;
  s.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), kBM, kLinBN);
  // few M tiles (projection head, small batches): split N across units so that more SMs take part
  if (s.num_units < sms && s.n_tiles > 1) {
    s.tiles_per_unit = 1;
    s.chunks = s.n_tiles;
    s.num_units = s.m_tiles * s.chunks;
  }
  StoreParams p;
  p.out = out;
  p.bias = bias;
  p.residual = residual;
  p.ldo = ldo;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.act = act;
  p.out_f32 = out_f32;
  p.fmt = fmt;
  using SM = GemmSmem<kLinBN, kLinStages>;
  auto kern = gemm_tc_kernel<EpiStore, kLinBN, kLinStages>;
  LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kDynamic));
  const int grid = s.num_units < sms ? s.num_units : sms;
  {
    KernelScope ks(kKcLinear, static_cast<cudaStream_t>(stream), 2.0 * M * static_cast<double>(N) * K,
                   (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 +
                       static_cast<double>(M) * N * (out_f32 ? 4.0 : 2.0) + (residual ? static_cast<double>(M) * N * 2.0 : 0.0));
    kern<<<grid, kGemmThreads, SM::kDynamic, static_cast<cudaStream_t>(stream)>>>(ta, tb, s, p);
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

}  // namespace ldot
