// Launch accounting for libldot_sm100a: every kernel launch of the library goes through a KernelScope, which
// (always) counts the launch per kernel class and (when ldot_prof_enable(1) was called) brackets it with a pair of
// CUDA events on the launching stream, so that bench.py can report per-kernel device time, achieved FLOP/s and
// bytes/s measured live inside its timed region (include/ldot.h: ldot_prof_*).
#pragma once
#include <cuda_runtime.h>

namespace ldot {

enum KernelClass : int {
  kKcCoarse = 0,     // tcgen05 score + top-k (coarse pass of the search)
  kKcSelect,         // candidate-list selection
  kKcRescore,        // exact fp32 rescoring + ranking + certificate
  kKcQueryPrep,      // fp32 -> 16-bit queries + statistics
  kKcIndexPrep,      // index centring / conversion
  kKcExactScan,      // exhaustive fallback
  kKcMerge,          // multi-shard top-k merge
  kKcLinear,         // tcgen05 encoder GEMM (+ bias / GELU / residual epilogue)
  kKcAttention,
  kKcLayerNorm,
  kKcEmbed,
  kKcCast,           // casts / hi-lo splits
  kKcNll,
  kKcOptim,          // AdamW / gradient norm
  kKcQkvAttn,        // fused Q|K|V projection + attention (tcgen05 main loop, mma.sync attention in the epilogue)
  kKcCount
};

const char* kernel_class_name(int c);
void prof_begin(int cls, cudaStream_t st, double flops, double bytes, void** token);
void prof_end(void* token, cudaStream_t st);

struct KernelScope {
  void* token;
  cudaStream_t st;
  KernelScope(int cls, cudaStream_t s, double flops = 0.0, double bytes = 0.0) : token(nullptr), st(s) {
    prof_begin(cls, s, flops, bytes, &token);
  }
  ~KernelScope() {
    if (token) prof_end(token, st);
  }
};

}  // namespace ldot
