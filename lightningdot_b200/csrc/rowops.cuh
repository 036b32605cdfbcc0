// Row-wise helpers shared by the tower kernels (encoder.cu) and the training kernels (train.cu): 16-bit <-> fp32
// vector loads / stores of 8 elements, warp reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace ldot {

constexpr float kLnEps = 1e-12f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

template <int FMT> __device__ __forceinline__ float2 cvt2(uint32_t u) {
  if (FMT == 1) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <int FMT> __device__ __forceinline__ uint32_t pk2(float a, float b) {
  if (FMT == 1) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int FMT> __device__ __forceinline__ void load8_16(const uint16_t* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 x = cvt2<FMT>(w[t]);
    f[2 * t] = x.x;
    f[2 * t + 1] = x.y;
  }
}
__device__ __forceinline__ void load8_f32(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <int FMT> __device__ __forceinline__ void store8_16(uint16_t* p, const float (&f)[8]) {
  uint4 u;
  u.x = pk2<FMT>(f[0], f[1]);
  u.y = pk2<FMT>(f[2], f[3]);
  u.z = pk2<FMT>(f[4], f[5]);
  u.w = pk2<FMT>(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

}  // namespace ldot
