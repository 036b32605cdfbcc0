// Counter-based dropout masks for the training step (dvl/models/bi_encoder.py:97-99: hidden_dropout_prob /
// attention_probs_dropout_prob of the towers).  The reference draws its masks from torch's Philox stream, which no other
// implementation can reproduce element for element; what has to match is the distribution (iid Bernoulli(1 - p) keeps,
// kept values scaled by 1 / (1 - p)) and that backward sees exactly the forward's mask.  Here a mask bit is a pure
// function of (seed of the forward call, site, element index), so the backward kernels regenerate it instead of storing
// it, and the CPU oracle restates the same function (oracle/dropout.py) to check values and gradients under dropout.
//
//   key  = mix32(seed_lo ^ mix32(seed_hi + 0x9E3779B9 * site))          (host, once per launch)
//   keep = mix32(idx_lo ^ mix32(idx_hi ^ key)) >= thr,  thr = floor(p * 2^32)
// Element index: hidden sites  row * cols + col  of the [tokens, H] activation; attention  ((b * heads + head) * S + i) * S + j.
#pragma once
#include <cstdint>

namespace ldot {

struct DropKey {
  uint32_t thr;      // 0 = dropout off
  uint32_t key;
  float inv_keep;
  const uint32_t* epoch;   // device word mixed into the key AT RUN TIME (null: none) - see ldot_dropout_epoch
};

// A captured CUDA graph replays its launches with the kernel arguments of capture time, so a host-side seed would
// repeat the same masks on every replay.  ldot_dropout_epoch(ptr) registers a device word that every later
// dropout-bearing launch reads when it RUNS: the replaying code bumps the word between replays and forward and backward
// of one replay see the same value.  Process-wide, like the stream a caller passes: set it around capture.
inline const uint32_t*& drop_epoch_slot() {
  static const uint32_t* slot = nullptr;
  return slot;
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {   // "lowbias32" integer finaliser
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

inline DropKey make_drop_key(float p, unsigned long long seed, int site) {
  DropKey k;
  k.epoch = nullptr;
  if (!(p > 0.f)) {
    k.thr = 0; k.key = 0; k.inv_keep = 1.f;
    return k;
  }
  k.epoch = drop_epoch_slot();
  const double t = static_cast<double>(p) * 4294967296.0;
  k.thr = t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
  k.key = mix32(static_cast<uint32_t>(seed) ^ mix32(static_cast<uint32_t>(seed >> 32) + 0x9E3779B9u * static_cast<uint32_t>(site)));
  k.inv_keep = 1.f / (1.f - p);
  return k;
}

#ifdef __CUDACC__
// first statement of every kernel that takes a DropKey
__device__ __forceinline__ DropKey drop_resolve(DropKey k) {
  if (k.epoch != nullptr) k.key = mix32(k.key ^ (0x85EBCA6Bu * __ldg(k.epoch)));
  return k;
}
#endif

__device__ __forceinline__ uint32_t drop_inner(const DropKey& k, unsigned long long idx) {
  return mix32(static_cast<uint32_t>(idx >> 32) ^ k.key);
}
__device__ __forceinline__ bool drop_keep(const DropKey& k, uint32_t inner, uint32_t idx_lo) {
  return mix32(idx_lo ^ inner) >= k.thr;
}
__device__ __forceinline__ bool drop_keep(const DropKey& k, unsigned long long idx) {
  return drop_keep(k, drop_inner(k, idx), static_cast<uint32_t>(idx));
}

}  // namespace ldot
