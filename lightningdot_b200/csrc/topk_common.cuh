// Shared device helpers for the top-k kernels: order-preserving float<->uint keys, packed (key, id) entries,
// warp- and block-level "k-th largest" selection by bisection on the key bits.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ldot {

// Monotone map fp32 -> u32: a < b  <=>  fkey(a) < fkey(b) (for non-NaN values; -0 < +0).
__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
  return __uint_as_float(u);
}

constexpr uint32_t kKeyNegInf = 0x007FFFFFu;  // fkey(-inf)

// A candidate entry: RAW fp32 bits of the coarse score in the high word, shard-local row id in the low word (the
// coarse epilogue appends entries with one predicated 8-byte store; consumers order them by entry_key()).
__device__ __forceinline__ unsigned long long pack_entry(float score, uint32_t id) {
  return (static_cast<unsigned long long>(__float_as_uint(score)) << 32) | id;
}
__device__ __forceinline__ uint32_t entry_key(unsigned long long e) {
  const uint32_t u = static_cast<uint32_t>(e >> 32);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Final ranking key: (exact fp32 score desc, row id asc)  ==  descending order of this 64-bit value.
__device__ __forceinline__ unsigned long long rank_key(float score, uint32_t id) {
  return (static_cast<unsigned long long>(fkey(score)) << 32) | (0xFFFFFFFFu - id);
}
__device__ __forceinline__ float rank_key_score(unsigned long long k) { return fkey_inv(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t rank_key_id(unsigned long long k) { return 0xFFFFFFFFu - static_cast<uint32_t>(k); }

// Block-wide sum of an int (all threads get the result).  `scratch` >= 33 ints of shared memory.
__device__ __forceinline__ int block_sum(int v, int* scratch) {
  v = __reduce_add_sync(0xFFFFFFFFu, v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int t = lane < nw ? scratch[lane] : 0;
    t = __reduce_add_sync(0xFFFFFFFFu, t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// Block-level: the k-th largest of n 64-bit values read through `get(i)` (values must be distinct or ties are
// acceptable as "any"); returns T such that count(v >= T) >= k and count(v > T) < k.  Requires n >= k >= 1.
template <class Get>
__device__ unsigned long long block_kth_largest_u64(Get get, int n, int k, int* scratch) {
  unsigned long long T = 0;
  for (int bit = 63; bit >= 0; --bit) {
    const unsigned long long cand = T | (1ull << bit);
    int c = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) c += (get(i) >= cand);
    if (block_sum(c, scratch) >= k) T = cand;
  }
  return T;
}

}  // namespace ldot
