// Exact inner-product top-k search (DenseFlatIndexer.search_knn, reference dvl/indexer/faiss_indexers.py:82-87,
// i.e. faiss.IndexFlatIP.search) as a B200 pipeline:
//
//   1. query_prepare   fp32 queries -> 16-bit copy + per-query norms / rounding residual / q.mu
//   2. coarse pass     tcgen05 GEMM  Q16[nq,d] x X16[n,d]^T  (fp32 accumulate in TMEM) whose epilogue keeps, per
//                      query, a thresholded candidate list - score rows never reach HBM
//   3. select          per query: the k' best coarse candidates over all index chunks + c_min (the k'-th coarse score)
//   4. rescore         exact scores (fp32 inputs, fp64 accumulation, rounded to fp32) of the k' candidates from the
//                      fp32 master index, ranked (score desc, row id asc); a per-query CERTIFICATE proves that no
//                      row outside the candidate list can belong to the top-k; uncertified queries are flagged
//   5. exact fallback  (separate entry point) full fp64-accumulated scan for flagged queries
//
// The 16-bit index copy is CENTRED (x - mean row): a per-query constant shift of all scores that leaves the
// ranking unchanged but removes the common component that dominates near-collinear embeddings.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cfloat>
#include <cmath>
#include "coarse_ts.cuh"
#include "gemm_tc.cuh"
#include "host_common.h"
#include "prof.h"
#include "search_plan.h"
#include "topk_common.cuh"

namespace ldot {

// ================================================================================================
// 2. coarse pass epilogue: thresholded candidate lists with warp-cooperative compaction
// ================================================================================================
struct TopKParams {
  int nq;                     // valid queries
  int n;                      // valid index rows
  int kprime;                 // entries kept by a compaction
  int cap;                    // list capacity per (unit, query) = EPL * 32  (>= kprime + 64)
  unsigned long long* cand;   // [num_units][128][cap] packed (coarse key, row id)
  int* cand_cnt;              // [num_units][128]
  unsigned int* gtau;         // [m_tiles * 128] best known lower bound of the k'-th coarse key, shared by all units
};

// The list of one lane is (nearly) full: keep its k' largest entries (in place), return the k'-th key.
// Whole-warp cooperative: entry i of the list is handled by lane i % 32.
template <int EPL>
__device__ __noinline__ uint32_t warp_compact(unsigned long long* buf, int n, int kprime, int lane, int& new_cnt) {
  constexpr bool kInRegs = EPL <= 16;  // small lists: one load round, entries stay in registers
  const unsigned lt = (1u << lane) - 1u;
  __syncwarp();
  uint32_t key[EPL];
  unsigned long long ent[kInRegs ? EPL : 1];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int idx = i * 32 + lane;
    const unsigned long long e = idx < n ? __ldcg(buf + idx) : 0ull;
    key[i] = idx < n ? entry_key(e) : 0u;
    if (kInRegs) ent[i] = e;
  }
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < EPL; ++i) c += (key[i] >= cand);
    if (__reduce_add_sync(0xFFFFFFFFu, c) >= kprime) T = cand;
  }
  int c = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) c += (key[i] > T);
  const int need_eq = kprime - __reduce_add_sync(0xFFFFFFFFu, c);
  int base = 0, eq_seen = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int idx = i * 32 + lane;
    const bool valid = idx < n;
    unsigned long long e;
    if (kInRegs) e = ent[i];
    else e = valid ? __ldcg(buf + idx) : 0ull;
    const bool is_eq = valid && key[i] == T;
    const unsigned m_eq = __ballot_sync(0xFFFFFFFFu, is_eq);
    const bool take = (key[i] > T) || (is_eq && (eq_seen + __popc(m_eq & lt)) < need_eq);
    const unsigned m_take = __ballot_sync(0xFFFFFFFFu, take);
    if (!kInRegs) __syncwarp();  // every lane has read chunk i before anyone overwrites positions <= i*32+31
    if (take) buf[base + __popc(m_take & lt)] = e;
    base += __popc(m_take);
    eq_seen += __popc(m_eq);
  }
  __syncwarp();
  new_cnt = base;
  return T;
}

template <int EPL, int BN>
struct EpiTopK {
  using Params = TopKParams;
  struct State {
    float tau;
    int cnt;
    int q;
    int tick;
    unsigned long long* buf;
  };

  static __device__ __forceinline__ void unit_begin(State& st, const Params& p, const UnitInfo& u, int row) {
    st.q = u.m_tile * kBM + row;
    st.buf = p.cand + (static_cast<size_t>(u.unit) * kBM + row) * p.cap;
    st.cnt = 0;
    st.tick = 0;
    st.tau = st.q < p.nq ? -INFINITY : INFINITY;  // padded query rows never collect anything
  }

  // Append entry (s, id) to the list at `ptr` when s > tau: one compare and three predicated instructions, no branch.
  static __device__ __forceinline__ void append_if(unsigned long long*& ptr, float s, float tau, uint32_t id) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.gt.f32 p, %1, %2;\n"
        "@p st.global.v2.b32 [%0], {%3, %4};\n"
        "@p add.u64 %0, %0, 8;\n"
        "}\n"
        : "+l"(ptr)
        : "f"(s), "f"(tau), "r"(id), "r"(__float_as_uint(s))
        : "memory");
  }

  static __device__ __forceinline__ void tile(State& st, const Params& p, const UnitInfo& u, int row, int nt,
                                              uint32_t taddr) {
    const int lane = threadIdx.x & 31;
    // the shared lower bound only moves when some list is compacted: poll it every 8th tile
    if (((st.tick++) & 7) == 0 && st.q < p.nq) {
      const uint32_t g = *reinterpret_cast<volatile unsigned int*>(p.gtau + st.q);
      if (g > kKeyNegInf) st.tau = fmaxf(st.tau, fkey_inv(g));
    }
    const int col0 = nt * BN;
    const bool full_tile = col0 + BN <= p.n;
#pragma unroll 1
    for (int c = 0; c < BN; c += 64) {
      // both 32-column loads of this 64-column slab are in flight before the wait
      uint32_t v0[32], v1[32];
      ptx::tmem_ld32(taddr + c, v0);
      ptx::tmem_ld32(taddr + c + 32, v1);
      ptx::tmem_ld_wait();
      if (full_tile) {
        unsigned long long* ptr = st.buf + st.cnt;
        const uint32_t id0 = static_cast<uint32_t>(col0 + c);
#pragma unroll
        for (int j = 0; j < 32; ++j) append_if(ptr, __uint_as_float(v0[j]), st.tau, id0 + j);
#pragma unroll
        for (int j = 0; j < 32; ++j) append_if(ptr, __uint_as_float(v1[j]), st.tau, id0 + 32 + j);
        st.cnt = static_cast<int>(ptr - st.buf);
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float s = __uint_as_float(j < 32 ? v0[j] : v1[j - 32]);
          if (s > st.tau && col0 + c + j < p.n) {
            st.buf[st.cnt] = pack_entry(s, static_cast<uint32_t>(col0 + c + j));
            ++st.cnt;
          }
        }
      }
      // a list that could overflow in the next 64 columns is compacted now, by the whole warp
      compact_where(st, p, lane, st.cnt > p.cap - 64);
    }
  }

  static __device__ __forceinline__ void compact_where(State& st, const Params& p, int lane, bool mine) {
    unsigned need = __ballot_sync(0xFFFFFFFFu, mine);
    while (need) {
      const int owner = __ffs(need) - 1;
      need &= need - 1;
      unsigned long long* obuf = reinterpret_cast<unsigned long long*>(
          __shfl_sync(0xFFFFFFFFu, reinterpret_cast<unsigned long long>(st.buf), owner));
      const int ocnt = __shfl_sync(0xFFFFFFFFu, st.cnt, owner);
      int new_cnt;
      const uint32_t T = warp_compact<EPL>(obuf, ocnt, p.kprime, lane, new_cnt);
      if (lane == owner) {
        st.cnt = new_cnt;
        st.tau = fmaxf(st.tau, fkey_inv(T));
        atomicMax(p.gtau + st.q, T);
      }
    }
  }

  static __device__ __forceinline__ void unit_end(State& st, const Params& p, const UnitInfo& u, int row) {
    // leave at most k' entries behind: bounds the union the select kernel has to look at (chunks x k')
    compact_where(st, p, threadIdx.x & 31, st.cnt > p.kprime);
    p.cand_cnt[static_cast<size_t>(u.unit) * kBM + row] = st.cnt;
  }
};

// ================================================================================================
// 2a. threshold pre-pass epilogue: the best score of every (query, sampled index tile)
// ================================================================================================
// The t-th largest of a query's tile maxima over a strided sample of S index tiles is a lower bound of its t-th best
// sample score (t different tiles each hold a row at least that good), hence a valid initial threshold for the main
// pass: it cuts the index down to ~ t * stride rows per query without any list traffic in the pre-pass itself.
struct TileMaxParams {
  int n;             // valid index rows
  int s_tiles;       // sampled tiles
  int tile_stride;   // sampled tile i is index tile i * tile_stride
  float* tmax;       // [m_tiles * 128][s_tiles]
};

template <int BN>
struct EpiTileMax {
  using Params = TileMaxParams;
  struct State {
    int q;
  };
  static __device__ __forceinline__ void unit_begin(State& st, const Params&, const UnitInfo& u, int row) {
    st.q = u.m_tile * kBM + row;
  }
  static __device__ __forceinline__ void unit_end(State&, const Params&, const UnitInfo&, int) {}
  static __device__ __forceinline__ void tile(State& st, const Params& p, const UnitInfo&, int, int nt, uint32_t taddr) {
    const int col0 = nt * BN;
    float m = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < BN; c += 64) {
      uint32_t v0[32], v1[32];
      ptx::tmem_ld32(taddr + c, v0);
      ptx::tmem_ld32(taddr + c + 32, v1);
      ptx::tmem_ld_wait();
      if (col0 + c + 64 <= p.n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (col0 + c + j < p.n) m = fmaxf(m, __uint_as_float(v0[j]));
          if (col0 + c + 32 + j < p.n) m = fmaxf(m, __uint_as_float(v1[j]));
        }
      }
    }
    p.tmax[static_cast<size_t>(st.q) * p.s_tiles + nt / p.tile_stride] = m;
  }
};

// one warp per query: gtau[q] = key of the t-th largest of its s_tiles tile maxima (0 = no threshold)
__global__ void __launch_bounds__(256) tau_kernel(const float* __restrict__ tmax, int nq, int s_tiles, int t,
                                                  unsigned int* __restrict__ gtau) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  const float* v = tmax + static_cast<size_t>(q) * s_tiles;
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
    for (int i = lane; i < s_tiles; i += 32) c += (fkey(__ldg(v + i)) >= cand);
    if (__reduce_add_sync(0xFFFFFFFFu, c) >= t) T = cand;
  }
  if (lane == 0) gtau[q] = (s_tiles >= t && T > kKeyNegInf) ? T : 0u;
}

// ================================================================================================
// 3. select: merge the per-chunk lists of one query into its k' best coarse candidates
// ================================================================================================
constexpr int kSelStage = 8192;    // entries staged in shared memory (64 KB)
constexpr int kSelGroup = 256;     // lists merged by one block (one list count per thread)

// Where the candidate lists of a query live.  List c of query q holds cnt[c * cnt_sc + q * cnt_sq] entries at
// ent + c * ent_sc + q * ent_sq.   Level 1 reads the per-unit lists of the coarse pass (c = index chunk), level 2
// reads the per-group survivors of level 1.
struct SelLists {
  const unsigned long long* ent;
  const int* cnt;
  long long ent_sc, ent_sq, cnt_sc, cnt_sq;
  int num_lists;
};

// grid (nq, groups).  Block (q, g) merges lists [g * kSelGroup, (g + 1) * kSelGroup) of query q into their k'
// largest coarse keys (bisection on the key bits over entries staged in shared memory, or streamed from the
// L2-resident lists when the union is larger).  final == 0: survivors (packed entries) go to out_ent / out_cnt for a
// second level; final == 1: survivor row ids go to sel_idx and the k'-th key to sel_cmin.
// Invariant kept at every level: an entry that is dropped has at least k' entries above it, so every row outside the
// final list has a coarse score <= c_min.
__global__ void __launch_bounds__(256) select_kernel(const SelLists in, const unsigned int* __restrict__ gtau, int kprime,
                                                     int final, unsigned long long* __restrict__ out_ent,
                                                     int* __restrict__ out_cnt, int* __restrict__ sel_idx,
                                                     float* __restrict__ sel_cmin, unsigned int* __restrict__ tau_out,
                                                     unsigned int* __restrict__ sel_key) {
  extern __shared__ unsigned long long sel_smem[];
  unsigned long long* stage = sel_smem;  // [kSelStage]
  __shared__ int cnts[kSelGroup];
  __shared__ int scratch[33];
  __shared__ int n_stage, out_pos, eq_pos;

  const int q = blockIdx.x, g = blockIdx.y, groups = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = g * kSelGroup;
  const int nl = in.num_lists - c0 < kSelGroup ? in.num_lists - c0 : kSelGroup;
  int mine = 0;
  if (threadIdx.x < kSelGroup) {
    const int v = threadIdx.x < nl ? in.cnt[(c0 + threadIdx.x) * in.cnt_sc + q * in.cnt_sq] : 0;
    cnts[threadIdx.x] = v;
    mine = v;
  }
  if (threadIdx.x == 0) n_stage = out_pos = eq_pos = 0;
  const int L = block_sum(mine, scratch);  // (its barriers also publish cnts and the counters)
  auto list_of = [&](int c) { return in.ent + (c0 + c) * in.ent_sc + q * in.ent_sq; };
  unsigned long long* oent = final ? nullptr : out_ent + (static_cast<size_t>(q) * groups + g) * kprime;
  int* oidx = final ? sel_idx + static_cast<size_t>(q) * kprime : nullptr;
  unsigned int* okey = (final && sel_key) ? sel_key + static_cast<size_t>(q) * kprime : nullptr;   // coarse keys, same order
  auto emit = [&](int pos, unsigned long long e) {
    if (final) {
      oidx[pos] = static_cast<int>(e & 0xFFFFFFFFu);
      if (okey) okey[pos] = entry_key(e);
    } else {
      oent[pos] = e;
    }
  };

  if (L < kprime) {
    // Fewer than k' entries: everything survives.  Rows that are in no list were dropped by a threshold, and every
    // threshold ever applied to this query is <= the final gtau[q] (atomicMax), so c_min = gtau[q]; gtau[q] == 0 means
    // that nothing was ever dropped (no compaction happened anywhere, no sampled threshold): c_min = -inf.
    for (int c = warp; c < nl; c += 8) {
      const unsigned long long* src = list_of(c);
      for (int i = lane; i < cnts[c]; i += 32) emit(atomicAdd(&out_pos, 1), __ldcg(src + i));
    }
    if (final) {
      __syncthreads();
      for (int i = L + threadIdx.x; i < kprime; i += blockDim.x) {
        oidx[i] = -1;
        if (okey) okey[i] = 0u;
      }
      if (threadIdx.x == 0) {
        const uint32_t g = gtau[q];
        sel_cmin[q] = g > kKeyNegInf ? fkey_inv(g) : -INFINITY;
        if (tau_out) tau_out[q] = 0u;  // sample too small to bound anything
      }
    } else if (threadIdx.x == 0) {
      out_cnt[q * groups + g] = L;
    }
    return;
  }
  // keys below the shared lower bound of the k'-th key cannot be among the k' best (at least k' entries of one
  // list are >= gtau, so the filter never leaves fewer than k' entries in the union that contains that list;
  // groups that do not contain it may shrink below k' - then they simply keep everything that is left)
  const uint32_t floor_key = gtau[q];
  const bool staged = L <= kSelStage;
  if (staged) {
    for (int c = warp; c < nl; c += 8) {
      const unsigned long long* src = list_of(c);
      for (int i = lane; i < cnts[c]; i += 32) {
        const unsigned long long e = __ldcg(src + i);
        if (entry_key(e) >= floor_key) stage[atomicAdd(&n_stage, 1)] = e;
      }
    }
    __syncthreads();
  }
  const int ns = n_stage;
  auto for_each = [&](auto&& f) {
    if (staged) {
      for (int i = threadIdx.x; i < ns; i += blockDim.x) f(stage[i]);
    } else {
      for (int c = warp; c < nl; c += 8) {
        const unsigned long long* src = list_of(c);
        for (int i = lane; i < cnts[c]; i += 32) f(__ldcg(src + i));
      }
    }
  };
  if (staged && ns < kprime) {
    for (int i = threadIdx.x; i < ns; i += blockDim.x) emit(i, stage[i]);
    // only possible in a non-final group that does not hold the list that set gtau: nothing >= gtau was dropped
    if (!final && threadIdx.x == 0) out_cnt[q * groups + g] = ns;
    if (final) {  // cannot happen (the union holds the list that set gtau); keep the output well-formed anyway
      for (int i = ns + threadIdx.x; i < kprime; i += blockDim.x) {
        oidx[i] = -1;
        if (okey) okey[i] = 0u;
      }
      if (threadIdx.x == 0) {
        sel_cmin[q] = fkey_inv(floor_key);
        if (tau_out) tau_out[q] = 0u;
      }
    }
    return;
  }
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t candk = T | (1u << bit);
    int c = 0;
    for_each([&](unsigned long long e) { c += (entry_key(e) >= candk); });
    if (block_sum(c, scratch) >= kprime) T = candk;
  }
  int c = 0;
  for_each([&](unsigned long long e) { c += (entry_key(e) > T); });
  const int n_gt = block_sum(c, scratch);
  const int need_eq = kprime - n_gt;
  for_each([&](unsigned long long e) {
    const uint32_t key = entry_key(e);
    if (key > T) {
      emit(atomicAdd(&out_pos, 1), e);
    } else if (key == T) {
      const int r = atomicAdd(&eq_pos, 1);
      if (r < need_eq) emit(n_gt + r, e);
    }
  });
  if (threadIdx.x == 0) {
    if (final) {
      sel_cmin[q] = fkey_inv(T);
      if (tau_out) tau_out[q] = T;
    } else {
      out_cnt[q * groups + g] = kprime;
    }
  }
}

// ================================================================================================
// 4. rescore + rank + certificate
// ================================================================================================
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// exact score of one (query, row) pair: fp32 products are exact in fp64, the 768-term sum carries ~1e-13 relative
// error, so the fp32 rounding of the result is (almost surely) the correctly rounded inner product.
__device__ __forceinline__ double warp_dot_f64(const float4* __restrict__ xr, const float4* qs, int d4, int lane) {
  double acc = 0.0;
  for (int i = lane; i < d4; i += 32) {
    const float4 a = __ldg(xr + i);
    const float4 b = qs[i];
    acc = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc);
    acc = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc);
    acc = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc);
    acc = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc);
  }
  return warp_sum_f64(acc);
}

__device__ void block_bitonic_desc(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], b = keys[ixj];
          const bool desc = (i & k) == 0;
          if (desc ? (a < b) : (a > b)) {
            keys[i] = b;
            keys[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

struct RescoreParams {
  const float* q;        // [nq, d] fp32 queries
  const float* x;        // [n, d] fp32 master index
  const int* sel_idx;    // [nq, kprime] candidate rows (-1 = none)
  const float* sel_cmin; // [nq]
  const unsigned int* sel_key;  // [nq, kprime] coarse keys of the candidates (same order as sel_idx)
  const float* tau;      // [nq] or null: a lower bound of the GLOBAL k-th best exact score (sharded search, phase 2)
  const float* qstats;   // [nq, 2]: |q16|, |q - q16|
  const double* qmu;     // [nq]: q . mu
  const float* xstats;   // [2]: max_j |x'_j - x16_j|, max_j |x16_j|
  float* out_scores;     // [nq, k]
  long long* out_idx;    // [nq, k]
  int* flags;            // [nq] 1 = not certified
  int* flag_count;
  long long id_offset;   // added to every returned row id (shard offset)
  int d, k, kprime, kp_pad;
};

// |exact centred score - coarse score| <= E: rounding of x, rounding of q (+ cross term), tensor-core fp32 accumulation
// (d additions, each within 2^-22 of the running sum of |products| <= |q16| |x16|: twice the bound of a truncating adder)
__device__ __forceinline__ double coarse_error_bound(const float* qstats, const float* xstats, int q, int d) {
  const double nq16 = qstats[2 * q], rq = qstats[2 * q + 1];
  const double rmax = xstats[0], xmax = xstats[1];
  return nq16 * rmax + rq * (xmax + rmax) + d * 2.384185791015625e-07 * nq16 * xmax;
}

// Sharded search, phase 1 -> phase 2 hand-off.  bound[q] = a value that at least m rows of THIS shard reach in exact
// score: (m-th best coarse score of the candidate list) + q.mu - E.  Over W shards with W m >= k, the minimum of the
// bounds is a lower bound of the global k-th best exact score (k rows reach it), so a shard only has to rescore - and
// only has to have kept - rows whose exact score can reach that minimum.  -inf when the list holds fewer than m rows.
__global__ void __launch_bounds__(256) shard_bound_kernel(const unsigned int* __restrict__ sel_key, const float* __restrict__ qstats,
                                                          const double* __restrict__ qmu, const float* __restrict__ xstats,
                                                          int kprime, int m, int d, float* __restrict__ bound) {
  extern __shared__ unsigned int sb_keys[];
  __shared__ unsigned int found;
  const int q = blockIdx.x;
  for (int i = threadIdx.x; i < kprime; i += blockDim.x) sb_keys[i] = sel_key[static_cast<size_t>(q) * kprime + i];
  if (threadIdx.x == 0) found = 0u;
  __syncthreads();
  // rank by counting: the entry with exactly m - 1 entries ahead of it (ties broken by position) is the m-th best
  for (int i = threadIdx.x; i < kprime; i += blockDim.x) {
    const unsigned int key = sb_keys[i];
    if (key == 0u) continue;
    int ahead = 0;
    for (int j = 0; j < kprime; ++j) ahead += (sb_keys[j] > key) || (sb_keys[j] == key && j < i);
    if (ahead == m - 1) found = key;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = -INFINITY;
    if (found > kKeyNegInf) {
      const double c = static_cast<double>(fkey_inv(found));
      const double E = coarse_error_bound(qstats, xstats, q, d);
      const double v = c + qmu[q] - E * 1.0001;
      b = static_cast<float>(v - 1.1920928955078125e-07 * (fabs(v) + 1e-30) - 1e-30);   // round towards -inf
      if (static_cast<double>(b) > v) b = nextafterf(b, -INFINITY);
    }
    bound[q] = b;
  }
}

__global__ void __launch_bounds__(1024) rescore_kernel(const RescoreParams p) {
  extern __shared__ unsigned long long rs_smem[];
  unsigned long long* keys = rs_smem;                              // [kp_pad]
  float4* qs = reinterpret_cast<float4*>(rs_smem + p.kp_pad);      // [d / 4]
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d4 = p.d >> 2;
  const float4* qrow = reinterpret_cast<const float4*>(p.q + static_cast<size_t>(q) * p.d);
  for (int i = threadIdx.x; i < d4; i += blockDim.x) qs[i] = qrow[i];
  for (int i = p.kprime + threadIdx.x; i < p.kp_pad; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  const int* cand = p.sel_idx + static_cast<size_t>(q) * p.kprime;
  const int nwarps = blockDim.x >> 5;
  // sharded search: a candidate whose exact score cannot reach the global lower bound tau is not rescored
  const bool pruned = p.tau != nullptr && p.tau[q] > -INFINITY;
  const double reach = pruned ? coarse_error_bound(p.qstats, p.xstats, q, p.d) * 1.0001 + p.qmu[q] : 0.0;
  const double tau = pruned ? static_cast<double>(p.tau[q]) : 0.0;
  for (int j = warp; j < p.kprime; j += nwarps) {
    const int id = cand[j];
    bool skip = id < 0;
    if (!skip && pruned) {
      const double up = static_cast<double>(fkey_inv(p.sel_key[static_cast<size_t>(q) * p.kprime + j])) + reach;
      skip = up + 1.1920928955078125e-07 * fabs(up) < tau;
    }
    if (skip) {
      if (lane == 0) keys[j] = 0ull;
      continue;
    }
    const double s = warp_dot_f64(reinterpret_cast<const float4*>(p.x + static_cast<size_t>(id) * p.d), qs, d4, lane);
    if (lane == 0) keys[j] = rank_key(static_cast<float>(s), static_cast<uint32_t>(id));
  }
  __syncthreads();
  block_bitonic_desc(keys, p.kp_pad);
  for (int r = threadIdx.x; r < p.k; r += blockDim.x) {
    const unsigned long long kk = keys[r];
    const bool ok = kk != 0ull;
    p.out_scores[static_cast<size_t>(q) * p.k + r] = ok ? rank_key_score(kk) : -FLT_MAX;
    p.out_idx[static_cast<size_t>(q) * p.k + r] = ok ? static_cast<long long>(rank_key_id(kk)) + p.id_offset : -1ll;
  }
  if (threadIdx.x == 0) {
    // Certificate.  Every row j outside the candidate list has coarse score c_j <= c_min, and its exact CENTRED score
    // s'_j = q.(x_j - mu) satisfies |s'_j - c_j| <= E, so s'_j <= c_min + E.  If the k-th best exact centred score
    // among the candidates is strictly larger, the true top-k is inside the list (and the list is ranked exactly).
    const float cmin = p.sel_cmin[q];
    bool certified;
    if (cmin == -INFINITY) {
      certified = true;  // no row was dropped anywhere
    } else if (pruned && static_cast<double>(cmin) + reach + 1.1920928955078125e-07 * fabs(static_cast<double>(cmin) + reach) < tau) {
      certified = true;  // no row outside the list can reach the global k-th best score: this shard's list is complete
    } else {
      const unsigned long long kk = keys[p.k - 1];
      if (kk == 0ull) {
        certified = false;
      } else {
        const double sk = static_cast<double>(rank_key_score(kk));
        const double E = coarse_error_bound(p.qstats, p.xstats, q, p.d);
        const double slack = 1.1920928955078125e-07 * (fabs(sk) + fabs(p.qmu[q])) + 1e-30;
        certified = (sk - p.qmu[q] - slack) > (static_cast<double>(cmin) + E * 1.0001);
      }
    }
    p.flags[q] = certified ? 0 : 1;
    if (!certified) atomicAdd(p.flag_count, 1);
  }
}

// ================================================================================================
// 1. query / index preparation
// ================================================================================================
template <class T> __device__ __forceinline__ T to16(float v);
template <> __device__ __forceinline__ __half to16<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 to16<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <class T> __device__ __forceinline__ float from16(T v);
template <> __device__ __forceinline__ float from16<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float from16<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// one warp per query row
template <class T>
__global__ void __launch_bounds__(256) query_prepare_kernel(const float* __restrict__ q, const float* __restrict__ mu,
                                                            int nq, int d, T* __restrict__ q16,
                                                            float* __restrict__ qstats, double* __restrict__ qmu) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= nq) return;
  const float* src = q + static_cast<size_t>(row) * d;
  T* dst = q16 + static_cast<size_t>(row) * d;
  double n16 = 0.0, r2 = 0.0, dm = 0.0;
  for (int i = lane; i < d; i += 32) {
    const float v = src[i];
    const T h = to16<T>(v);
    dst[i] = h;
    const double hv = static_cast<double>(from16<T>(h));
    n16 = fma(hv, hv, n16);
    const double e = static_cast<double>(v) - hv;
    r2 = fma(e, e, r2);
    dm = fma(static_cast<double>(v), static_cast<double>(mu[i]), dm);
  }
  n16 = warp_sum_f64(n16);
  r2 = warp_sum_f64(r2);
  dm = warp_sum_f64(dm);
  if (lane == 0) {
    qstats[2 * row] = static_cast<float>(sqrt(n16) * 1.000001);
    qstats[2 * row + 1] = static_cast<float>(sqrt(r2) * 1.000001) + FLT_MIN;
    qmu[row] = dm;
  }
}

// column sums in fp64 (thread per column, coalesced over columns; rows strided over blocks)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long n, int d,
                                                     double* __restrict__ acc) {
  const long long rows_per_block = (n + gridDim.x - 1) / gridDim.x;
  const long long r0 = blockIdx.x * rows_per_block;
  const long long r1 = min(n, r0 + rows_per_block);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double s = 0.0;
    for (long long r = r0; r < r1; ++r) s += static_cast<double>(x[r * d + c]);
    atomicAdd(acc + c, s);
  }
}

__global__ void mean_finalize_kernel(const double* __restrict__ acc, long long n, int d, int center,
                                     float* __restrict__ mu) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) mu[c] = center ? static_cast<float>(acc[c] / static_cast<double>(n)) : 0.0f;
}

// one warp per index row: x16 = T(fl32(x - mu)); residual against the exact (x - mu); global maxima
template <class T>
__global__ void __launch_bounds__(256) index_convert_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                                            long long n, int d, T* __restrict__ x16,
                                                            unsigned int* __restrict__ stats_bits) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const float* src = x + row * d;
  T* dst = x16 + row * d;
  double r2 = 0.0, n2 = 0.0;
  for (int i = lane; i < d; i += 32) {
    const float v = src[i], m = mu[i];
    const T h = to16<T>(v - m);
    dst[i] = h;
    const double hv = static_cast<double>(from16<T>(h));
    const double e = (static_cast<double>(v) - static_cast<double>(m)) - hv;
    r2 = fma(e, e, r2);
    n2 = fma(hv, hv, n2);
  }
  r2 = warp_sum_f64(r2);
  n2 = warp_sum_f64(n2);
  if (lane == 0) {
    // non-negative floats order like their bit patterns
    atomicMax(stats_bits + 0, __float_as_uint(static_cast<float>(sqrt(r2) * 1.000001) + FLT_MIN));
    atomicMax(stats_bits + 1, __float_as_uint(static_cast<float>(sqrt(n2) * 1.000001) + FLT_MIN));
  }
}

// ================================================================================================
// 5. exact fallback: full scan with fp64 accumulation for a small batch of queries
// ================================================================================================
constexpr int kExactBatch = 8;

__global__ void __launch_bounds__(256) exact_scores_kernel(const float* __restrict__ q, int nf, const float* __restrict__ x,
                                                           long long n, int d, float* __restrict__ scores) {
  extern __shared__ float4 ex_q[];  // [nf][d/4]
  const int d4 = d >> 2;
  for (int i = threadIdx.x; i < nf * d4; i += blockDim.x) ex_q[i] = reinterpret_cast<const float4*>(q)[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < n; row += static_cast<long long>(gridDim.x) * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * d);
    double acc[kExactBatch];
#pragma unroll
    for (int f = 0; f < kExactBatch; ++f) acc[f] = 0.0;
    for (int i = lane; i < d4; i += 32) {
      const float4 a = __ldg(xr + i);
#pragma unroll
      for (int f = 0; f < kExactBatch; ++f) {
        if (f < nf) {
          const float4 b = ex_q[f * d4 + i];
          acc[f] = fma(static_cast<double>(a.x), static_cast<double>(b.x), acc[f]);
          acc[f] = fma(static_cast<double>(a.y), static_cast<double>(b.y), acc[f]);
          acc[f] = fma(static_cast<double>(a.z), static_cast<double>(b.z), acc[f]);
          acc[f] = fma(static_cast<double>(a.w), static_cast<double>(b.w), acc[f]);
        }
      }
    }
#pragma unroll
    for (int f = 0; f < kExactBatch; ++f) {
      if (f < nf) {
        const double s = warp_sum_f64(acc[f]);
        if (lane == 0) scores[static_cast<size_t>(f) * n + row] = static_cast<float>(s);
      }
    }
  }
}

// Level 1 of the exact selection: grid (chunks, queries); the block ranks one chunk of a query's score row in shared
// memory and emits its k best (rank keys: score desc, row id asc; keys are distinct) to out_keys[f][chunk][0..k),
// zero padded.
constexpr int kExactChunk = 6144;  // scores per block: 48 KB of 64-bit rank keys

__global__ void __launch_bounds__(256) exact_chunk_topk_kernel(const float* __restrict__ scores, long long n, int k,
                                                               unsigned long long* __restrict__ out_keys) {
  extern __shared__ unsigned long long ec_keys[];  // [kExactChunk]
  __shared__ int scratch[33];
  __shared__ int out_pos;
  const int f = blockIdx.y, c = blockIdx.x, chunks = gridDim.x;
  const long long off = static_cast<long long>(c) * kExactChunk;
  const int len = static_cast<int>(n - off < kExactChunk ? n - off : kExactChunk);
  const float* s = scores + static_cast<size_t>(f) * n + off;
  unsigned long long* out = out_keys + (static_cast<size_t>(f) * chunks + c) * k;
  if (threadIdx.x == 0) out_pos = 0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) ec_keys[i] = rank_key(__ldcg(s + i), static_cast<uint32_t>(off + i));
  __syncthreads();
  if (len <= k) {
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = i < len ? ec_keys[i] : 0ull;
    return;
  }
  auto get = [&](int i) { return ec_keys[i]; };
  const unsigned long long T = block_kth_largest_u64(get, len, k, scratch);
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const unsigned long long e = ec_keys[i];
    if (e >= T) out[atomicAdd(&out_pos, 1)] = e;   // exactly k entries: keys are distinct
  }
}

// Level 2: one block per query over the m = chunks * k level-1 keys -> the final ranked top-k.
__global__ void __launch_bounds__(1024) exact_merge_kernel(const unsigned long long* __restrict__ keys, int m, long long n,
                                                           int k, int kp_pad, long long id_offset,
                                                           float* __restrict__ out_scores, long long* __restrict__ out_idx) {
  extern __shared__ unsigned long long es_keys[];  // [kp_pad]
  __shared__ int scratch[33];
  __shared__ int out_pos;
  const int f = blockIdx.x;
  const unsigned long long* src = keys + static_cast<size_t>(f) * m;
  const int kk = static_cast<long long>(k) < n ? k : static_cast<int>(n);
  if (threadIdx.x == 0) out_pos = 0;
  for (int i = threadIdx.x; i < kp_pad; i += blockDim.x) es_keys[i] = 0ull;
  __syncthreads();
  auto get = [&](int i) { return __ldcg(src + i); };
  const unsigned long long T = block_kth_largest_u64(get, m, kk, scratch);
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const unsigned long long e = get(i);
    if (e >= T && e != 0ull) es_keys[atomicAdd(&out_pos, 1)] = e;
  }
  __syncthreads();
  block_bitonic_desc(es_keys, kp_pad);
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const unsigned long long e = es_keys[r];
    const bool ok = e != 0ull;
    out_scores[static_cast<size_t>(f) * k + r] = ok ? rank_key_score(e) : -FLT_MAX;
    out_idx[static_cast<size_t>(f) * k + r] = ok ? static_cast<long long>(rank_key_id(e)) + id_offset : -1ll;
  }
}

// ================================================================================================
// multi-GPU: merge W shard results (already exact, global ids) into the global top-k, same ranking rule
// ================================================================================================
// Every shard list arrives RANKED ((score desc, id asc); label -1 entries at the end), and keys are unique (global ids),
// so the global rank of entry r of list w is r + sum over the other lists of how many of their entries beat it:
// W - 1 binary searches per entry in shared memory and one synchronisation, instead of a bitonic sort of all W k keys.
// ws_s / ws_i: elements between the lists of consecutive shards (nq * k when the two arrays are dense [W, nq, k]; larger when
// every shard's scores and ids arrive in ONE packed buffer, sharded.py).
__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ scores, const long long* __restrict__ idx,
                                                    int W, int nq, int k, long long ws_s, long long ws_i,
                                                    float* __restrict__ out_scores, long long* __restrict__ out_idx) {
  extern __shared__ unsigned long long mg_keys[];  // [W * k]
  const int q = blockIdx.x;
  const int total = W * k;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int w = i / k, r = i - w * k;
    const size_t o = static_cast<size_t>(q) * k + r;
    const long long id = idx[static_cast<size_t>(w) * ws_i + o];
    mg_keys[i] = id >= 0 ? rank_key(scores[static_cast<size_t>(w) * ws_s + o], static_cast<uint32_t>(id)) : 0ull;
  }
  for (int r = threadIdx.x; r < k; r += blockDim.x) {   // (fewer than k valid entries in total: the tail stays empty)
    out_scores[static_cast<size_t>(q) * k + r] = -FLT_MAX;
    out_idx[static_cast<size_t>(q) * k + r] = -1ll;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const unsigned long long e = mg_keys[i];
    if (e == 0ull) continue;
    const int w = i / k;
    int pos = i - w * k;
    for (int w2 = 0; w2 < W && pos < k; ++w2) {
      if (w2 == w) continue;
      const unsigned long long* l = mg_keys + w2 * k;
      int lo = 0, hi = k;   // first position of list w2 whose key is <= e  (= number of its entries that beat e)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (l[mid] > e) lo = mid + 1;
        else hi = mid;
      }
      pos += lo;
    }
    if (pos < k) {
      out_scores[static_cast<size_t>(q) * k + pos] = rank_key_score(e);
      out_idx[static_cast<size_t>(q) * k + pos] = static_cast<long long>(rank_key_id(e));
    }
  }
}

// ================================================================================================
// host side
// ================================================================================================
static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int search_make_plan(SearchPlan* pl, long long nq, long long n, int d, int k, int coarse_k, int sms) {
  LDOT_REQUIRE(nq >= 1 && nq <= (1 << 22), "nq=%lld out of range [1, 4194304] (split the query batch)", nq);
  LDOT_REQUIRE(n >= 1 && n <= 2000000000ll, "n=%lld out of range", n);
  LDOT_REQUIRE(d >= 8 && d <= 4096 && d % 8 == 0, "d=%d must be a multiple of 8 in [8, 4096]", d);
  LDOT_REQUIRE(k >= 1 && k <= 1024, "k=%d out of range [1, 1024]", k);
  int kp = coarse_k > 0 ? coarse_k : k + (k / 2 > 32 ? k / 2 : 32);
  if (coarse_k <= 0 && kp > 1280) kp = 1280;  // k close to the 1024 limit: whatever margin is left
  if (kp < 64) kp = 64;
  kp = (kp + 31) / 32 * 32;
  LDOT_REQUIRE(kp >= k && kp <= 1280, "coarse_k=%d must be in [k, 1280]", kp);
  pl->kprime = kp;
  // list capacity cap = epl * 32 >= k' + 64 (a list is compacted to its k' best when fewer than 64 slots are left)
  pl->epl = kp <= 192 ? 8 : kp <= 448 ? 16 : kp <= 960 ? 32 : 80;
  pl->cap = pl->epl * 32;
  pl->kp_pad = next_pow2(kp);
  // queries parked in TMEM (A-stationary kernel) whenever the vector fits 384 TMEM columns
  pl->a_in_tmem = d <= kTsMaxD ? 1 : 0;
  pl->bn = pl->a_in_tmem ? kTsBN : kSearchBN;
  pl->m_tiles = static_cast<int>((nq + kBM - 1) / kBM);
  pl->n_tiles = static_cast<int>((n + pl->bn - 1) / pl->bn);
  // Units = m_tiles x chunks, dealt round-robin to one persistent CTA per SM.  Pick the chunk count that fills the
  // last wave best; prefer long chunks (>= 4096 index rows: fewer candidate lists, fewer list compactions) and, on
  // ties, fewer chunks.
  int min_tpu = 4096 / pl->bn;
  // ... unless the index (shard) is so small that long chunks would leave SMs without a unit: 125 k rows in 4096-row
  // chunks are 28 units on 148 SMs (81 us for the 128-query pass, measured); down to 512-row chunks when units are scarce
  {
    const long long units_long = static_cast<long long>(pl->m_tiles) * (pl->n_tiles / min_tpu > 0 ? pl->n_tiles / min_tpu : 1);
    if (units_long < sms) {
      int t = static_cast<int>((static_cast<long long>(pl->m_tiles) * pl->n_tiles + sms - 1) / sms);
      const int floor_tpu = 512 / pl->bn > 0 ? 512 / pl->bn : 1;
      if (t < floor_tpu) t = floor_tpu;
      if (t < min_tpu) min_tpu = t;
    }
  }
  int max_chunks = pl->n_tiles / min_tpu;
  if (max_chunks < 1) max_chunks = 1;
  if (max_chunks > 2048) max_chunks = 2048;
  // With several query tiles the CTAs that run together share index chunks through L2: keep two chunks (the ones in
  // flight at any time) well inside the 126 MB L2.
  int c_min = 1;
  if (pl->m_tiles > 1) {
    const long long tile_bytes = static_cast<long long>(pl->bn) * d * 2;
    long long max_tpu = (24ll << 20) / tile_bytes;
    if (max_tpu < 8) max_tpu = 8;
    c_min = static_cast<int>((pl->n_tiles + max_tpu - 1) / max_tpu);
    if (c_min > max_chunks) max_chunks = c_min;
  }
  int best_chunks = c_min;
  double best_eff = -1.0;
  for (int c = c_min; c <= max_chunks; ++c) {
    const int tpu_c = (pl->n_tiles + c - 1) / c;
    const int chunks_c = (pl->n_tiles + tpu_c - 1) / tpu_c;
    const long long units = static_cast<long long>(pl->m_tiles) * chunks_c;
    const long long waves = (units + sms - 1) / sms;
    // cost of the busiest CTA in tiles vs the perfectly balanced share
    const double eff = (static_cast<double>(pl->m_tiles) * pl->n_tiles / sms) / (static_cast<double>(waves) * tpu_c);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best_chunks = chunks_c;
    }
    if (units >= 32ll * sms && c > c_min) break;
  }
  pl->tiles_per_unit = (pl->n_tiles + best_chunks - 1) / best_chunks;
  pl->chunks = (pl->n_tiles + pl->tiles_per_unit - 1) / pl->tiles_per_unit;
  pl->num_units = pl->m_tiles * pl->chunks;
  pl->groups = (pl->chunks + kSelGroup - 1) / kSelGroup;  // > 1: two-level select
  LDOT_REQUIRE(pl->groups <= kSelGroup, "index too large for one search call (chunks=%d)", pl->chunks);
  // Threshold pre-pass.  Every s_stride-th index tile is scored first and the s_kprime-th best sample score of a
  // query becomes its initial threshold.  With sample fraction f = 1 / s_stride the threshold cuts the full index
  // down to ~ s_kprime * s_stride rows per query (>= 3 k'), and it is too high - fewer than k' rows above it, which
  // the select stage detects and flags - only if >= s_kprime of the k' best rows fell into the sample:
  // P(Binomial(k', f) >= 32) with k' f <= 10, below 1e-8.
  pl->sample = 0;
  pl->s_kprime = 32;
  pl->s_stride = pl->s_tiles = pl->s_tiles_per_unit = pl->s_chunks = pl->s_units = 0;
  {
    int stride = kp / 2;
    if (stride > pl->n_tiles / 64) stride = pl->n_tiles / 64;  // at least 64 sample tiles (4096 rows)
    // (skipping the pre-pass for small shards was measured: 125 k rows x 128 queries 59 -> 782 us - without an initial
    // threshold every list is compacted several times per unit)
    if (stride >= 2 && stride * 10 >= kp) {
      pl->sample = 1;
      pl->s_stride = stride;
      pl->s_tiles = (pl->n_tiles + stride - 1) / stride;
      int sc = (sms + pl->m_tiles - 1) / pl->m_tiles;   // the pre-pass keeps no lists: chunks only spread the work
      if (sc > pl->s_tiles) sc = pl->s_tiles;
      pl->s_tiles_per_unit = (pl->s_tiles + sc - 1) / sc;
      pl->s_chunks = (pl->s_tiles + pl->s_tiles_per_unit - 1) / pl->s_tiles_per_unit;
      pl->s_units = pl->m_tiles * pl->s_chunks;
    }
  }
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  pl->off_q16 = take(static_cast<size_t>(nq) * d * 2);
  pl->off_qstats = take(static_cast<size_t>(nq) * 2 * sizeof(float));
  pl->off_qmu = take(static_cast<size_t>(nq) * sizeof(double));
  pl->off_gtau = take(static_cast<size_t>(pl->m_tiles) * kBM * sizeof(unsigned int));
  pl->off_flagcnt = take(sizeof(int));
  pl->off_tmax = take(static_cast<size_t>(pl->m_tiles) * kBM * sizeof(float) * (pl->sample ? pl->s_tiles : 0));
  pl->off_cnt = take(static_cast<size_t>(pl->num_units) * kBM * sizeof(int));
  pl->off_sel_idx = take(static_cast<size_t>(nq) * kp * sizeof(int));
  pl->off_sel_cmin = take(static_cast<size_t>(nq) * sizeof(float));
  pl->off_sel_key = take(static_cast<size_t>(nq) * kp * sizeof(unsigned int));
  pl->off_l2_ent = take(pl->groups > 1 ? static_cast<size_t>(nq) * pl->groups * kp * sizeof(unsigned long long) : 0);
  pl->off_l2_cnt = take(pl->groups > 1 ? static_cast<size_t>(nq) * pl->groups * sizeof(int) : 0);
  pl->off_cand = take(static_cast<size_t>(pl->num_units) * kBM * pl->cap * sizeof(unsigned long long));
  pl->total_bytes = off;
  return kOk;
}

template <int EPL>
static int launch_coarse_ss(const CUtensorMap& ta, const CUtensorMap& tb, const GemmSched& s, const TopKParams& p, int sms,
                            cudaStream_t st) {
  using SM = GemmSmem<kSearchBN, kSearchStages>;
  auto kern = gemm_tc_kernel<EpiTopK<EPL, kSearchBN>, kSearchBN, kSearchStages>;
  static bool configured = false;  // (one device per process)
  if (!configured) {
    LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kDynamic));
    configured = true;
  }
  const int grid = s.num_units < sms ? s.num_units : sms;
  kern<<<grid, kGemmThreads, SM::kDynamic, st>>>(ta, tb, s, p);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

template <int EPL>
static int launch_coarse_ts(const CUtensorMap& tb, const CUtensorMap* tb3, const GemmSched& s, const TsQueries& tq,
                            const TopKParams& p, int sms, cudaStream_t st) {
  const int grid = s.num_units < sms ? s.num_units : sms;
  if (tb3) {  // d == 768: compile-time unrolled TMA / MMA issue loops
    auto kern = coarse_ts_kernel<EpiTopK<EPL, kTsBN>, kTsFastKb>;
    static bool configured = false;  // (one device per process)
    if (!configured) {
      LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TsSmem::kDynamic));
      configured = true;
    }
    kern<<<grid, kGemmThreads, TsSmem::kDynamic, st>>>(tb, *tb3, s, tq, p);
  } else {
    auto kern = coarse_ts_kernel<EpiTopK<EPL, kTsBN>, 0>;
    static bool configured = false;
    if (!configured) {
      LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TsSmem::kDynamic));
      configured = true;
    }
    kern<<<grid, kGemmThreads, TsSmem::kDynamic, st>>>(tb, tb, s, tq, p);
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

static int launch_tilemax_ss(const CUtensorMap& ta, const CUtensorMap& tb, const GemmSched& s, const TileMaxParams& p,
                             int sms, cudaStream_t st) {
  using SM = GemmSmem<kSearchBN, kSearchStages>;
  auto kern = gemm_tc_kernel<EpiTileMax<kSearchBN>, kSearchBN, kSearchStages>;
  static bool configured = false;
  if (!configured) {
    LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kDynamic));
    configured = true;
  }
  const int grid = s.num_units < sms ? s.num_units : sms;
  kern<<<grid, kGemmThreads, SM::kDynamic, st>>>(ta, tb, s, p);
  LDOT_CHECK_LAUNCH();
  return kOk;
}

static int launch_tilemax_ts(const CUtensorMap& tb, const CUtensorMap* tb3, const GemmSched& s, const TsQueries& tq,
                             const TileMaxParams& p, int sms, cudaStream_t st) {
  const int grid = s.num_units < sms ? s.num_units : sms;
  if (tb3) {
    auto kern = coarse_ts_kernel<EpiTileMax<kTsBN>, kTsFastKb>;
    static bool configured = false;
    if (!configured) {
      LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TsSmem::kDynamic));
      configured = true;
    }
    kern<<<grid, kGemmThreads, TsSmem::kDynamic, st>>>(tb, *tb3, s, tq, p);
  } else {
    auto kern = coarse_ts_kernel<EpiTileMax<kTsBN>, 0>;
    static bool configured = false;
    if (!configured) {
      LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TsSmem::kDynamic));
      configured = true;
    }
    kern<<<grid, kGemmThreads, TsSmem::kDynamic, st>>>(tb, tb, s, tq, p);
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

int search_run(const SearchArgs& a) {
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  SearchPlan pl;
  if (int e = search_make_plan(&pl, a.nq, a.n, a.d, a.k, a.coarse_k, sms)) return e;
  LDOT_REQUIRE(a.ws_bytes >= pl.total_bytes, "workspace too small: %zu < %zu", a.ws_bytes, pl.total_bytes);
  LDOT_REQUIRE(a.coarse_dtype == 0 || a.coarse_dtype == 1, "coarse_dtype must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(a.x16) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.ws) & 255) == 0,
               "q, x, x16 must be 16-byte aligned and the workspace 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(a.stream);
  uint8_t* ws = static_cast<uint8_t*>(a.ws);
  void* q16 = ws + pl.off_q16;
  float* qstats = reinterpret_cast<float*>(ws + pl.off_qstats);
  double* qmu = reinterpret_cast<double*>(ws + pl.off_qmu);
  unsigned int* gtau = reinterpret_cast<unsigned int*>(ws + pl.off_gtau);
  int* flagcnt = reinterpret_cast<int*>(ws + pl.off_flagcnt);
  int* cnt = reinterpret_cast<int*>(ws + pl.off_cnt);
  int* sel_idx = reinterpret_cast<int*>(ws + pl.off_sel_idx);
  float* sel_cmin = reinterpret_cast<float*>(ws + pl.off_sel_cmin);
  unsigned int* sel_key = reinterpret_cast<unsigned int*>(ws + pl.off_sel_key);
  unsigned long long* l2_ent = reinterpret_cast<unsigned long long*>(ws + pl.off_l2_ent);
  int* l2_cnt = reinterpret_cast<int*>(ws + pl.off_l2_cnt);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + pl.off_cand);
  const int nq = static_cast<int>(a.nq);

  // exact rescoring of the selected candidates + ranking + certificate (phase 2 of a sharded search starts here)
  auto rescore_stage = [&]() -> int {
  RescoreParams rp;
    rp.q = a.q;
    rp.x = a.x;
    rp.sel_idx = sel_idx;
    rp.sel_cmin = sel_cmin;
    rp.sel_key = sel_key;
    rp.tau = a.tau;
    rp.qstats = qstats;
    rp.qmu = qmu;
    rp.xstats = a.xstats;
    rp.out_scores = a.out_scores;
    rp.out_idx = a.out_idx;
    rp.flags = a.out_flags;
    rp.flag_count = flagcnt;
    rp.id_offset = a.id_offset;
    rp.d = a.d;
    rp.k = a.k;
    rp.kprime = pl.kprime;
    rp.kp_pad = pl.kp_pad;
    const size_t rs_smem = pl.kp_pad * sizeof(unsigned long long) + static_cast<size_t>(a.d) * sizeof(float);
    {
      KernelScope ks(kKcRescore, st, 2.0 * nq * pl.kprime * a.d, static_cast<double>(nq) * pl.kprime * a.d * 4.0);
      // one block per query; each warp gathers its candidates' fp32 rows one after the other, so a small batch (online
      // regime: fewer blocks than the machine has room for) gets 32 warps per block - 5 dependent row gathers instead of 20
      rescore_kernel<<<nq, nq <= 4 * sms ? 1024 : 256, rs_smem, st>>>(rp);
    }
    LDOT_CHECK_LAUNCH();
    if (a.out_flag_count)
      LDOT_CUDA(cudaMemcpyAsync(a.out_flag_count, flagcnt, sizeof(int), cudaMemcpyDefault, st));  // device or pinned host
    return kOk;
  };
  if (a.phase == 2) return rescore_stage();

  // gtau and the flag counter are adjacent in the plan: one memset clears both
  float* tmax = reinterpret_cast<float*>(ws + pl.off_tmax);
  LDOT_CUDA(cudaMemsetAsync(gtau, 0, (pl.off_flagcnt - pl.off_gtau) + sizeof(int), st));
  const int qblocks = (nq + 7) / 8;
  {
    KernelScope ks(kKcQueryPrep, st, 0.0, static_cast<double>(nq) * a.d * 6.0);
    if (a.coarse_dtype == 0)
      query_prepare_kernel<__half><<<qblocks, 256, 0, st>>>(a.q, a.mu, nq, a.d, static_cast<__half*>(q16), qstats, qmu);
    else
      query_prepare_kernel<__nv_bfloat16><<<qblocks, 256, 0, st>>>(a.q, a.mu, nq, a.d,
                                                                   static_cast<__nv_bfloat16*>(q16), qstats, qmu);
  }
  LDOT_CHECK_LAUNCH();

  CUtensorMap ta, tb;
  if (int e = make_tmap_kmajor_16b(&tb, a.x16, a.n, a.d, static_cast<uint64_t>(a.d) * 2, pl.bn)) return e;
  if (!pl.a_in_tmem)
    if (int e = make_tmap_kmajor_16b(&ta, q16, a.nq, a.d, static_cast<uint64_t>(a.d) * 2, kBM)) return e;
  // d == 768: 3-D view of the index for the compile-time unrolled kernel (one TMA instruction per 4 K blocks)
  CUtensorMap tb3;
  const CUtensorMap* tb3p = nullptr;
  if (pl.a_in_tmem && a.d == kTsFastKb * kBK) {
    if (int e = make_tmap_kblocks_16b(&tb3, a.x16, a.n, a.d, static_cast<uint64_t>(a.d) * 2, kTsBN, kTsKbPerStage)) return e;
    tb3p = &tb3;
  }
  TsQueries tq;
  tq.q16 = static_cast<const uint16_t*>(q16);
  tq.nq = nq;
  tq.d = a.d;

  // one coarse pass (tcgen05 score + per-list top-k') over the tiles of schedule `s`
  auto run_coarse = [&](const GemmSched& s, const TopKParams& tp, int epl, double rows) -> int {
    // algorithmic work: every (query, row) pair once; the 16-bit rows and queries read once
    KernelScope coarse_scope(kKcCoarse, st, 2.0 * nq * rows * a.d, (rows + nq) * a.d * 2.0);
    if (pl.a_in_tmem) {
      switch (epl) {
        case 8: return launch_coarse_ts<8>(tb, tb3p, s, tq, tp, sms, st);
        case 16: return launch_coarse_ts<16>(tb, tb3p, s, tq, tp, sms, st);
        case 32: return launch_coarse_ts<32>(tb, tb3p, s, tq, tp, sms, st);
        default: return launch_coarse_ts<80>(tb, tb3p, s, tq, tp, sms, st);
      }
    }
    switch (epl) {
      case 8: return launch_coarse_ss<8>(ta, tb, s, tp, sms, st);
      case 16: return launch_coarse_ss<16>(ta, tb, s, tp, sms, st);
      case 32: return launch_coarse_ss<32>(ta, tb, s, tp, sms, st);
      default: return launch_coarse_ss<80>(ta, tb, s, tp, sms, st);
    }
  };
  // where the lists of a coarse pass live: list (chunk c, query q) = unit (c * m_tiles + q / 128), row q % 128
  auto lists_of = [&](int chunks, int cap) {
    SelLists l;
    l.ent = cand;
    l.cnt = cnt;
    l.ent_sc = static_cast<long long>(pl.m_tiles) * kBM * cap;
    l.ent_sq = cap;
    l.cnt_sc = static_cast<long long>(pl.m_tiles) * kBM;
    l.cnt_sq = 1;
    l.num_lists = chunks;
    return l;
  };
  const size_t sel_smem = kSelStage * sizeof(unsigned long long);
  static_assert(kSelStage * sizeof(unsigned long long) <= 64 * 1024, "select staging");
  {
    static bool configured = false;
    if (!configured) {
      LDOT_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sel_smem)));
      configured = true;
    }
  }

  GemmSched s;
  s.m_tiles = pl.m_tiles;
  s.k_blocks = (a.d + kBK - 1) / kBK;
  s.idesc = ptx::make_idesc_f16(a.coarse_dtype == 0 ? 0u : 1u, kBM, pl.bn);
  TopKParams tp;
  tp.nq = nq;
  tp.n = static_cast<int>(a.n);
  tp.cand = cand;
  tp.cand_cnt = cnt;

  if (pl.sample) {
    // threshold pre-pass: tile maxima over a strided sample of the index tiles -> initial gtau[q]
    s.n_tiles = pl.s_tiles;
    s.tiles_per_unit = pl.s_tiles_per_unit;
    s.chunks = pl.s_chunks;
    s.num_units = pl.s_units;
    s.tile_stride = pl.s_stride;
    TileMaxParams mp;
    mp.n = static_cast<int>(a.n);
    mp.s_tiles = pl.s_tiles;
    mp.tile_stride = pl.s_stride;
    mp.tmax = tmax;
    double s_rows = static_cast<double>(pl.s_tiles) * pl.bn;
    if (s_rows > static_cast<double>(a.n)) s_rows = static_cast<double>(a.n);
    {
      // (time is charged to the coarse class, FLOPs / bytes are not: the sampled rows are scored again by the main pass, so
      // they are overhead of THIS algorithm, not algorithmic work of the search - bench.py's roofline_search)
      (void)s_rows;
      KernelScope ks(kKcCoarse, st, 0.0, 0.0);
      const int e = pl.a_in_tmem ? launch_tilemax_ts(tb, tb3p, s, tq, mp, sms, st) : launch_tilemax_ss(ta, tb, s, mp, sms, st);
      if (e) return e;
    }
    {
      KernelScope ks(kKcSelect, st);
      tau_kernel<<<(nq + 7) / 8, 256, 0, st>>>(tmax, nq, pl.s_tiles, pl.s_kprime, gtau);
    }
    LDOT_CHECK_LAUNCH();
  }

  s.n_tiles = pl.n_tiles;
  s.tiles_per_unit = pl.tiles_per_unit;
  s.chunks = pl.chunks;
  s.num_units = pl.num_units;
  s.tile_stride = 1;
  tp.kprime = pl.kprime;
  tp.cap = pl.cap;
  tp.gtau = gtau;
  if (int e = run_coarse(s, tp, pl.epl, static_cast<double>(a.n))) return e;

  const SelLists l1 = lists_of(pl.chunks, pl.cap);
  if (pl.groups == 1) {
    KernelScope ks(kKcSelect, st);
    select_kernel<<<dim3(nq, 1), 256, sel_smem, st>>>(l1, gtau, pl.kprime, 1, nullptr, nullptr, sel_idx, sel_cmin, nullptr,
                                                      sel_key);
    LDOT_CHECK_LAUNCH();
  } else {
    {
      KernelScope ks(kKcSelect, st);
      select_kernel<<<dim3(nq, pl.groups), 256, sel_smem, st>>>(l1, gtau, pl.kprime, 0, l2_ent, l2_cnt, nullptr, nullptr,
                                                                nullptr, nullptr);
    }
    LDOT_CHECK_LAUNCH();
    SelLists l2;
    l2.ent = l2_ent;
    l2.cnt = l2_cnt;
    l2.ent_sc = pl.kprime;
    l2.ent_sq = static_cast<long long>(pl.groups) * pl.kprime;
    l2.cnt_sc = 1;
    l2.cnt_sq = pl.groups;
    l2.num_lists = pl.groups;
    {
      KernelScope ks(kKcSelect, st);
      select_kernel<<<dim3(nq, 1), 256, sel_smem, st>>>(l2, gtau, pl.kprime, 1, nullptr, nullptr, sel_idx, sel_cmin, nullptr,
                                                        sel_key);
    }
    LDOT_CHECK_LAUNCH();
  }

  if (a.phase == 1) {   // sharded search: stop here and report what this shard can guarantee (see shard_bound_kernel)
    LDOT_REQUIRE(a.bound_out != nullptr && a.bound_m >= 1 && a.bound_m <= pl.kprime, "phase 1 needs bound_out and 1 <= m <= k'");
    KernelScope ks(kKcSelect, st);
    shard_bound_kernel<<<nq, 256, pl.kprime * sizeof(unsigned int), st>>>(sel_key, qstats, qmu, a.xstats, pl.kprime, a.bound_m,
                                                                          a.d, a.bound_out);
    LDOT_CHECK_LAUNCH();
    return kOk;
  }
  return rescore_stage();
}

size_t index_prepare_workspace_bytes(int d) { return align_up(static_cast<size_t>(d) * sizeof(double), 256); }

int index_prepare_run(const float* x, long long n, int d, int coarse_dtype, int center, void* x16, float* mu,
                      float* xstats, void* ws, size_t ws_bytes, void* stream) {
  LDOT_REQUIRE(n >= 1 && d >= 8 && d % 8 == 0, "bad index shape n=%lld d=%d", n, d);
  LDOT_REQUIRE(coarse_dtype == 0 || coarse_dtype == 1, "coarse_dtype must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(ws_bytes >= index_prepare_workspace_bytes(d), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* acc = static_cast<double*>(ws);
  LDOT_CUDA(cudaMemsetAsync(acc, 0, static_cast<size_t>(d) * sizeof(double), st));
  LDOT_CUDA(cudaMemsetAsync(xstats, 0, 2 * sizeof(float), st));
  if (center) {
    int blocks = static_cast<int>(n < 4096 ? (n + 7) / 8 : 1184);
    if (blocks < 1) blocks = 1;
    {
      KernelScope ks(kKcIndexPrep, st, 0.0, static_cast<double>(n) * d * 4.0);
      colsum_kernel<<<blocks, 256, 0, st>>>(x, n, d, acc);
    }
    LDOT_CHECK_LAUNCH();
  }
  {
    KernelScope ks(kKcIndexPrep, st);
    mean_finalize_kernel<<<(d + 255) / 256, 256, 0, st>>>(acc, n, d, center, mu);
  }
  LDOT_CHECK_LAUNCH();
  const int blocks = static_cast<int>((n + 7) / 8);
  KernelScope ks_conv(kKcIndexPrep, st, 0.0, static_cast<double>(n) * d * 6.0);
  if (coarse_dtype == 0)
    index_convert_kernel<__half><<<blocks, 256, 0, st>>>(x, mu, n, d, static_cast<__half*>(x16),
                                                         reinterpret_cast<unsigned int*>(xstats));
  else
    index_convert_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(x, mu, n, d, static_cast<__nv_bfloat16*>(x16),
                                                                reinterpret_cast<unsigned int*>(xstats));
  LDOT_CHECK_LAUNCH();
  return kOk;
}

static size_t exact_scores_bytes(long long n) { return align_up(static_cast<size_t>(kExactBatch) * n * sizeof(float), 256); }
static int exact_chunks(long long n) { return static_cast<int>((n + kExactChunk - 1) / kExactChunk); }

size_t exact_workspace_bytes(long long n) {
  // score rows of one query batch + level-1 keys sized for the largest k (1024)
  return exact_scores_bytes(n) + align_up(static_cast<size_t>(kExactBatch) * exact_chunks(n) * 1024 * sizeof(unsigned long long), 256);
}

int exact_run(const float* q, long long nf, const float* x, long long n, int d, int k, long long id_offset,
              float* out_scores, long long* out_idx, void* ws, size_t ws_bytes, void* stream) {
  LDOT_REQUIRE(nf >= 0 && n >= 1 && n <= 2000000000ll && d >= 8 && d % 8 == 0 && d <= 4096, "bad shape");
  LDOT_REQUIRE(k >= 1 && k <= 1024, "k=%d out of range [1, 1024]", k);
  LDOT_REQUIRE(ws_bytes >= exact_workspace_bytes(n), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  float* scores = static_cast<float*>(ws);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ws) + exact_scores_bytes(n));
  const int chunks = exact_chunks(n);
  const int kp_pad = next_pow2(k);
  LDOT_CUDA(cudaFuncSetAttribute(exact_chunk_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(kExactChunk * sizeof(unsigned long long))));
  const size_t q_smem = static_cast<size_t>(kExactBatch) * d * sizeof(float);
  LDOT_CUDA(cudaFuncSetAttribute(exact_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(q_smem)));
  for (long long f0 = 0; f0 < nf; f0 += kExactBatch) {
    const int nb = static_cast<int>(nf - f0 < kExactBatch ? nf - f0 : kExactBatch);
    long long blocks = (n + 7) / 8;
    if (blocks > sms * 8) blocks = sms * 8;
    {
      KernelScope ks(kKcExactScan, st, 2.0 * nb * static_cast<double>(n) * d, static_cast<double>(n) * d * 4.0);
      exact_scores_kernel<<<static_cast<int>(blocks), 256, static_cast<size_t>(nb) * d * sizeof(float), st>>>(
          q + f0 * d, nb, x, n, d, scores);
    }
    LDOT_CHECK_LAUNCH();
    {
      KernelScope ks(kKcExactScan, st, 0.0, static_cast<double>(nb) * n * 4.0);
      exact_chunk_topk_kernel<<<dim3(chunks, nb), 256, kExactChunk * sizeof(unsigned long long), st>>>(scores, n, k, keys);
    }
    LDOT_CHECK_LAUNCH();
    {
      KernelScope ks(kKcExactScan, st, 0.0, static_cast<double>(nb) * chunks * k * 8.0);
      exact_merge_kernel<<<nb, 1024, kp_pad * sizeof(unsigned long long), st>>>(
          keys, chunks * k, n, k, kp_pad, id_offset, out_scores + f0 * k, out_idx + f0 * k);
    }
    LDOT_CHECK_LAUNCH();
  }
  return kOk;
}

int merge_run(const float* scores, const long long* idx, int W, long long nq, int k, long long ws_s, long long ws_i,
              float* out_scores, long long* out_idx, void* stream) {
  if (ws_s <= 0) ws_s = nq * k;
  if (ws_i <= 0) ws_i = nq * k;
  LDOT_REQUIRE(W >= 1 && W <= 64 && nq >= 0 && k >= 1 && k <= 1024, "bad merge shape W=%d nq=%lld k=%d", W, nq, k);
  if (nq == 0) return kOk;
  const size_t smem = static_cast<size_t>(W) * k * sizeof(unsigned long long);
  LDOT_REQUIRE(smem <= 200 * 1024, "W*k=%d too large for the merge kernel", W * k);
  LDOT_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  {
    KernelScope ks(kKcMerge, static_cast<cudaStream_t>(stream), 0.0, static_cast<double>(nq) * k * 12.0 * (W + 1));
    merge_kernel<<<static_cast<int>(nq), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        scores, idx, W, static_cast<int>(nq), k, ws_s, ws_i, out_scores, out_idx);
  }
  LDOT_CHECK_LAUNCH();
  return kOk;
}

}  // namespace ldot
