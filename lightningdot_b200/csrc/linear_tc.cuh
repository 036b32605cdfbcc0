// Encoder Linear on tcgen05:  out[M, N] = act(A[M, K] . W[N, K]^T + bias) (+ residual)
//
// Persistent, warp-specialised, one CTA per SM, 320 threads:
//   warp 0      TMA producer: A (activations) 128 x 64 and W 256 x 64 K-major boxes, SWIZZLE_128B, 4-stage ring
//   warp 1      TMEM owner + tcgen05.mma issuer (128 x 256 x 16 per instruction, fp32 accumulate, 2 accumulators)
//   warps 2..   epilogue (8 warps; 16 for the GELU form): warp w reads TMEM lane quarter (w % 4) and every
//               ((w - 2) / 4)-th 32-column box of the 128 x 256 tile in
//               16-column slices, software-pipelined: tcgen05.ld of slice s + 1 is in flight while slice s gets bias,
//               erf-GELU, residual in registers | swizzled st.shared into the warp's staging buffer; every two
//               slices one TMA store (cp.async.bulk.tensor) of the 32 x 32 box.  Stores are coalesced by the TMA unit and clipped at the
//               M / N edges of the output, so there is no tail handling on the store side.
// Tiles are visited n-fastest (tile t = m_tile * n_tiles + n_tile): CTAs that run at the same time share the A row
// block (the big operand - activations) through L2 while the weights stay L2-resident anyway.
//
// CTAS = 2 (large M): the kernel runs as CTA PAIRS (cluster of 2, tcgen05 cta_group::2).  One pair owns a 256 x 256
// output tile; each CTA stages its own 128 rows of A and HALF of the W tile (128 of the 256 weight rows), the leader
// CTA's elected thread issues one 256 x 256 x 16 MMA for both SMs, and each CTA's epilogue drains its own 128 x 256
// half of the accumulator from its own TMEM.  Per SM this halves the W bytes pulled from L2 and written to shared
// memory (32 KB instead of 48 KB per 64-wide K block), which is what bounds the single-CTA form: 148 SMs x 96 B/clk
// of operand traffic is above what the L2 delivers (~43 B/clk/SM).  The smaller stage also buys a 6-deep ring.
//   full[s]    lives in the LEADER: both CTAs' TMA loads complete_tx on it, the leader's producer posts the expect_tx
//   empty[s]   one per CTA, signalled by the leader's multicast tcgen05.commit
//   tfull[a]   one per CTA, same multicast commit after the last K block
//   tempty[a]  lives in the LEADER: 2 x 8 epilogue warps arrive (the peer's remotely, release.cluster)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "gemm_tc.cuh"
#include "gelu.cuh"
#include "dropout.cuh"

namespace ldot {

constexpr int kLinBN = 256;
constexpr int kLinStages = 4;
constexpr int kLinEpiWarps = 8;
constexpr int kLinThreads = 32 * (2 + kLinEpiWarps);
// The erf-GELU epilogue (FFN-up: K = 768 behind 3072 output columns) is paced by per-warp instruction latency, not by
// a pipe: it runs with 16 epilogue warps (4 per scheduler, 96 registers each) and pays for their staging buffers with
// two pipeline stages.  Measured at 320k tokens: 8 warps 879, 12 warps 1004, 16 warps 1076 TFLOP/s.  The GELU' form of
// the backward (ACT = 2, heavier per element, 128 registers) runs with 12.
// The training form of the GELU epilogue (TRAIN: also stores the pre-activation and rounds it before GELU) does not fit 96
// registers - ncu showed local-memory reloads on its critical path, tensor pipe 39 % - and runs with 12 warps of 128.
__host__ __device__ constexpr int lin_epi_warps(int act, int train = 0) {
  return act == 1 ? (train ? 12 : 16) : act == 2 ? 12 : 8;   // (act 3: out = acc * aux, as light as the residual add)
}
__host__ __device__ constexpr int lin_threads(int act, int train = 0) { return 32 * (2 + lin_epi_warps(act, train)); }

struct LinSched {
  int m_tiles, n_tiles, num_tiles, k_blocks;
  uint32_t idesc;
  // split-K (RED kernels only): tile t = (k_split * m_tiles + m_tile) * n_tiles + n_tile covers K blocks
  // [k_split * kb_per_split, ...); num_tiles = m_tiles * n_tiles * k_splits.  1 / k_blocks otherwise.
  int k_splits, kb_per_split;
};

struct LinParams {
  const float* bias;      // [N] or null
  const void* residual;   // [M, ldr] 16-bit of `fmt`, or null
  long long ldr;
  long long M;
  int N;
  int fmt;                // 0 = fp16, 1 = bf16 (A, W, residual, 16-bit output)
  // training forward (ACT 0): out = dropout(acc + bias) (+ residual), mask element index row * N + col (drop.thr 0: off)
  DropKey drop;
  // training forward (ACT 1): GELU'(z) of the pre-activation z = acc + bias is ALSO written, 16-bit [M, ld_pre] (null: not
  // written) - the next GEMM needs GELU(z), backward only GELU'(z)
  void* pre;
  long long ld_pre;
};

template <int CTAS, int EW = kLinEpiWarps>
struct LinSmemT {
  static constexpr int kStages = (CTAS == 2 ? 6 : kLinStages) - (EW > 12 ? 2 : EW > 8 ? 1 : 0);
  static constexpr int kBRows = kLinBN / CTAS;         // W rows staged by one CTA
  static constexpr int kABytes = kBM * kBK * 2;        // 16 KB
  static constexpr int kBBytes = kBRows * kBK * 2;     // 32 KB (16 KB per CTA of a pair)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingPerWarp = 4096;         // one 32 x 32 fp32 box, or two 32 x 32 16-bit boxes
  static constexpr int kStagingOffset = kStages * kStageBytes;
  static constexpr int kBarOffset = kStagingOffset + EW * kStagingPerWarp;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;       // slack for manual 1024 B alignment
};
using LinSmem = LinSmemT<1>;
static_assert(LinSmemT<1>::kDynamic <= 227 * 1024 && LinSmemT<2>::kDynamic <= 227 * 1024 &&
              LinSmemT<1, 16>::kDynamic <= 227 * 1024 && LinSmemT<2, 16>::kDynamic <= 227 * 1024 &&
              LinSmemT<1, 12>::kDynamic <= 227 * 1024 && LinSmemT<2, 12>::kDynamic <= 227 * 1024, "linear kernel shared memory");

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

// sched.m_tiles counts 128 * CTAS-row tile rows; sched.num_tiles = m_tiles * n_tiles tiles of (128 * CTAS) x 256.
//
// Backward-pass forms of the same kernel (training):
//   AMN / BMN = 1   the operand is MN-major: the global matrix is [K, M] (resp. [K, N]) row-major and is used without
//                   a transpose pass - dgrad reads the nn.Linear weight [out, in] as B[N = in, K = out], wgrad reads
//                   dY [tokens, out] as A[M = out, K = tokens] and X [tokens, in] as B[N = in, K = tokens].  TMA lands
//                   64 (K rows) x 64 (MN, 128 B) SWIZZLE_128B boxes, 8 KB apart per 64-wide MN block; the shared-memory
//                   descriptor carries LBO = 8192 (MN block pitch), SBO = 1024 (8 K rows), and one 16-deep K step
//                   advances the start address by 2048 B.
//   ACT = 2         out = acc * gelu'(aux[m, n]) with aux = p.residual (a saved FFN-up pre-activation)
//   ACT = 3         out = acc * aux[m, n] (aux = the GELU'(z) the training forward stored)
//   RED = 1         fp32 output boxes are ADDED to global memory (cp.reduce.async.bulk.tensor .add): gradient
//                   accumulation, and what makes split-K (sched.k_splits > 1) a pure scheduling decision.
//   TRAIN = 1       the training-forward epilogues: dropout of the dense output (ACT 0, p.drop) / the pre-activation
//                   written next to GELU's output (ACT 1, p.pre).  A template flag so that the inference
//                   instantiations keep their register budget (the 16-warp GELU form has 96 per thread).
template <int ACT, int OUT_F32, int CTAS, int AMN = 0, int BMN = 0, int RED = 0, int TRAIN = 0>
__global__ void __launch_bounds__(lin_threads(ACT, TRAIN), 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_out, const LinSched sched, const LinParams p) {
  constexpr int EW = lin_epi_warps(ACT, TRAIN);   // epilogue warps
  constexpr int EC = EW / 4;               // ... per TMEM lane quarter: they share the row block's 32-column boxes
  using SM = LinSmemT<CTAS, EW>;
  constexpr int kStages = SM::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * SM::kABytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;   // 0 = leader of the pair
  const int first_tile = static_cast<int>(blockIdx.x) / CTAS;
  const int tile_step = static_cast<int>(gridDim.x) / CTAS;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], EW * CTAS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_out);
  }
  if (warp == 1) {
    if (CTAS == 2) {
      ptx::tmem_alloc_2cta(tmem_ptr, 512);
      ptx::tmem_relinquish_2cta();
    } else {
      ptx::tmem_alloc(tmem_ptr, 512);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) ptx::cluster_sync_all();  // the peer's barriers must be initialised before anything is signalled on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs of a pair)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const int mn_tiles = sched.m_tiles * sched.n_tiles;
      for (int t = first_tile; t < sched.num_tiles; t += tile_step) {
        const int ks = t / mn_tiles, tt = t - ks * mn_tiles;
        const int m_tile = tt / sched.n_tiles, n_tile = tt - m_tile * sched.n_tiles;
        const int a_row = (m_tile * CTAS + rank) * kBM;
        const int w_row = n_tile * kLinBN + rank * SM::kBRows;
        const int kb0 = ks * sched.kb_per_split;
        const int kb1 = kb0 + sched.kb_per_split < sched.k_blocks ? kb0 + sched.kb_per_split : sched.k_blocks;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * SM::kABytes;
          uint8_t* sb = smem_b + stage * SM::kBBytes;
          if (CTAS == 2) {
            // bytes of BOTH CTAs are credited to the leader's barrier
            const uint32_t lbar = ptx::mapa(ptx::smem_u32(&full[stage]), 0);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * SM::kStageBytes);
            if (AMN) {
#pragma unroll
              for (int j = 0; j < kBM / 64; ++j)
                ptx::tma_load_2d_2cta(sa + j * 8192, &tmap_a, lbar, a_row + j * 64, kb * kBK, ptx::kEvictNormal);
            } else {
              ptx::tma_load_2d_2cta(sa, &tmap_a, lbar, kb * kBK, a_row, ptx::kEvictNormal);
            }
            if (BMN) {
#pragma unroll
              for (int j = 0; j < SM::kBRows / 64; ++j)
                ptx::tma_load_2d_2cta(sb + j * 8192, &tmap_w, lbar, w_row + j * 64, kb * kBK, ptx::kEvictLast);
            } else {
              ptx::tma_load_2d_2cta(sb, &tmap_w, lbar, kb * kBK, w_row, ptx::kEvictLast);
            }
          } else {
            ptx::mbar_arrive_expect_tx(&full[stage], SM::kStageBytes);
            if (AMN) {
#pragma unroll
              for (int j = 0; j < kBM / 64; ++j)
                ptx::tma_load_2d(sa + j * 8192, &tmap_a, &full[stage], a_row + j * 64, kb * kBK, ptx::kEvictNormal);
            } else {
              ptx::tma_load_2d(sa, &tmap_a, &full[stage], kb * kBK, a_row, ptx::kEvictNormal);
            }
            if (BMN) {
#pragma unroll
              for (int j = 0; j < SM::kBRows / 64; ++j)
                ptx::tma_load_2d(sb + j * 8192, &tmap_w, &full[stage], w_row + j * 64, kb * kBK, ptx::kEvictLast);
            } else {
              ptx::tma_load_2d(sb, &tmap_w, &full[stage], kb * kBK, w_row, ptx::kEvictLast);
            }
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (the leader CTA only)
    if (rank == 0 && ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int mn_tiles = sched.m_tiles * sched.n_tiles;
      // descriptor start-address step (in 16 B units) of one 16-deep K slice: 32 B inside the 128 B row of a K-major
      // tile, two 8-row groups (2048 B) of an MN-major one
      constexpr uint64_t kAStep = AMN ? 128 : 2, kBStep = BMN ? 128 : 2;
      for (int t = first_tile; t < sched.num_tiles; t += tile_step) {
        const int ks = t / mn_tiles;
        const int kb0 = ks * sched.kb_per_split;
        const int kb1 = kb0 + sched.kb_per_split < sched.k_blocks ? kb0 + sched.kb_per_split : sched.k_blocks;
        // (the epilogue warps hand the accumulator back with CTA-scope arrives: only TMEM reads are ordered, by
        // tcgen05 fences - a cluster-scope acquire here would invalidate L1 on every poll of the spin)
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kLinBN);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_a + stage * SM::kABytes);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * SM::kBBytes);
          const uint64_t adesc = AMN ? ptx::make_smem_desc_sw128_mn(sa) : ptx::make_smem_desc_sw128(sa);
          const uint64_t bdesc = BMN ? ptx::make_smem_desc_sw128_mn(sb) : ptx::make_smem_desc_sw128(sb);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t acc = (kb != kb0 || k != 0) ? 1u : 0u;
            if (CTAS == 2) ptx::mma_f16_ss_2cta(tmem_d, adesc + kAStep * k, bdesc + kBStep * k, sched.idesc, acc);
            else ptx::mma_f16_ss(tmem_d, adesc + kAStep * k, bdesc + kBStep * k, sched.idesc, acc);
          }
          if (CTAS == 2) ptx::mma_commit_2cta(&empty[stage], 0x3);
          else ptx::mma_commit(&empty[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CTAS == 2) ptx::mma_commit_2cta(&tfull[as], 0x3);
        else ptx::mma_commit(&tfull[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (EW warps)
    const int quarter = warp & 3;
    const int cgrp = (warp - 2) >> 2;   // boxes cgrp, cgrp + EC, ... of every tile row block belong to this warp
    const int row = quarter * 32 + lane;
    uint8_t* staging = smem + SM::kStagingOffset + (warp - 2) * SM::kStagingPerWarp;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int as = 0;
    uint32_t aphase = 0;
    uint32_t nstore = 0;
    const DropKey drop = TRAIN ? drop_resolve(p.drop) : p.drop;
    // tempty of the LEADER collects the arrivals of both CTAs' epilogue warps
    const uint32_t tempty_addr[2] = {CTAS == 2 ? ptx::mapa(ptx::smem_u32(&tempty[0]), 0) : ptx::smem_u32(&tempty[0]),
                                     CTAS == 2 ? ptx::mapa(ptx::smem_u32(&tempty[1]), 0) : ptx::smem_u32(&tempty[1])};
    auto release_acc = [&](int a) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) ptx::mbar_arrive_remote(tempty_addr[a]);
        else ptx::mbar_arrive(&tempty[a]);
      }
    };
    const int mn_tiles = sched.m_tiles * sched.n_tiles;
    // The residual / aux operand of the epilogue (p.residual: one 32-byte piece per thread and slice, rows 2 N bytes apart)
    // is requested into L2 a whole tile ahead: read on first use it costs a DRAM round trip per slice on the epilogue's
    // critical path (the GELU' dgrad of the training step: tensor pipe 38 %, every sample on the first use of the load).
    auto prefetch_aux = [&](int t) {
      if (ACT == 1 || p.residual == nullptr || t >= sched.num_tiles) return;   // (the GELU forms take no residual)
      const int tt = t % mn_tiles;
      const int m_pair = tt / sched.n_tiles, n_tile = tt - m_pair * sched.n_tiles;
      const long long grow = static_cast<long long>(m_pair * CTAS + rank) * kBM + row;
      if (grow >= p.M) return;
      const uint16_t* base = static_cast<const uint16_t*>(p.residual) + grow * p.ldr + n_tile * kLinBN;
      for (int bx = cgrp; bx < kLinBN / 32; bx += EC)
        if (n_tile * kLinBN + bx * 32 < p.N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + bx * 32));
    };
    prefetch_aux(first_tile);
    for (int t = first_tile; t < sched.num_tiles; t += tile_step) {
      const int ks = t / mn_tiles, tt = t - ks * mn_tiles;
      const int m_pair = tt / sched.n_tiles, n_tile = tt - m_pair * sched.n_tiles;
      const int m_tile = m_pair * CTAS + rank;   // 128-row block of this CTA
      const float* bias = ks == 0 ? p.bias : nullptr;   // (split-K: the first slice carries the bias)
      prefetch_aux(t + tile_step);
      const long long grow = static_cast<long long>(m_tile) * kBM + row;
      const bool row_ok = grow < p.M;
      const int tile_col = n_tile * kLinBN;
      int nboxes = (p.N - tile_col + 31) / 32;   // 32-column boxes of this tile that hold valid columns
      nboxes = nboxes > kLinBN / 32 ? kLinBN / 32 : nboxes;
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = lane_base + static_cast<uint32_t>(as * kLinBN);
      if (cgrp >= nboxes) {
        release_acc(as);
      } else {
        // TMEM loads are software-pipelined over 16-column slices: slice s + 1 is in flight while slice s goes through
        // the elementwise tail.  Two slices fill one 32 x 32 staging box, which leaves as one TMA store.
        uint32_t v2[2][16];
        ptx::tmem_ld16(taddr + cgrp * 32, v2[0]);
        for (int bx = cgrp; bx < nboxes; bx += EC) {
          uint8_t* buf = OUT_F32 ? staging : staging + (nstore & 1u) * (SM::kStagingPerWarp / 2);
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            const int col = tile_col + bx * 32 + hs * 16;
            const bool valid = col < p.N;               // (warp-uniform)
            const bool full_slice = col + 16 <= p.N;
            uint32_t (&v)[16] = v2[hs];
            // operands of the elementwise tail are fetched while the TMEM load is in flight
            float4 b4[4];
            if (bias != nullptr && full_slice) {
#pragma unroll
              for (int j = 0; j < 4; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(bias + col) + j);
            }
            uint4 r4[2];
            const bool res_vec = p.residual != nullptr && row_ok && full_slice;
            if (res_vec) {
              const uint4* r = reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + grow * p.ldr + col);
              r4[0] = __ldg(r);
              r4[1] = __ldg(r + 1);
            }
            ptx::tmem_ld_wait();
            if (hs == 0) ptx::tmem_ld16(taddr + bx * 32 + 16, v2[1]);
            else if (bx + EC < nboxes) ptx::tmem_ld16(taddr + (bx + EC) * 32, v2[0]);
            else release_acc(as);   // accumulator drained: hand it back to the MMA warp before the math
            if (hs == 0) {          // the staging buffer of the store two boxes back must have been read
              if (lane == 0) {
                if (OUT_F32) ptx::bulk_wait_group_read<0>();
                else ptx::bulk_wait_group_read<1>();
              }
              __syncwarp();
            }
            if (valid) {
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
              if (bias != nullptr) {
                if (full_slice) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    f[4 * j] += b4[j].x;
                    f[4 * j + 1] += b4[j].y;
                    f[4 * j + 2] += b4[j].z;
                    f[4 * j + 3] += b4[j].w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (col + j < p.N) f[j] += __ldg(bias + col + j);
                }
              }
              if (TRAIN && ACT == 0 && drop.thr != 0) {
                // hidden-state dropout of the dense output (layer.py:113,154); 8 consecutive indices share a high word
                const unsigned long long idx = static_cast<unsigned long long>(grow) * p.N + col;
#pragma unroll
                for (int g8 = 0; g8 < 2; ++g8) {
                  const uint32_t inner = drop_inner(drop, idx + 8 * g8);
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    f[g8 * 8 + j] = drop_keep(drop, inner, static_cast<uint32_t>(idx) + g8 * 8 + j) ? f[g8 * 8 + j] * drop.inv_keep : 0.f;
                }
              }
              if (ACT == 1) {
                if (TRAIN) {
                  // training forward: GELU of the 16-bit-ROUNDED pre-activation (the tensor amp hands to F.gelu in the
                  // reference) and, from the same exponential, GELU'(z) - stored as 16 bit in p.pre: all that backward
                  // needs of z, so that the dgrad epilogue is one multiply instead of a second GELU evaluation
                  float gp[16];
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float2 z = unpack2(pack2(f[2 * j], f[2 * j + 1], p.fmt), p.fmt);
                    gelu_erf_both(z.x, f[2 * j], gp[2 * j]);
                    gelu_erf_both(z.y, f[2 * j + 1], gp[2 * j + 1]);
                  }
                  if (p.pre != nullptr && row_ok) {
                    uint16_t* pre = static_cast<uint16_t*>(p.pre) + grow * p.ld_pre + col;
                    if (full_slice) {
#pragma unroll
                      for (int j = 0; j < 2; ++j) {
                        uint4 w;
                        w.x = pack2(gp[8 * j], gp[8 * j + 1], p.fmt);
                        w.y = pack2(gp[8 * j + 2], gp[8 * j + 3], p.fmt);
                        w.z = pack2(gp[8 * j + 4], gp[8 * j + 5], p.fmt);
                        w.w = pack2(gp[8 * j + 6], gp[8 * j + 7], p.fmt);
                        reinterpret_cast<uint4*>(pre)[j] = w;
                      }
                    } else {
#pragma unroll
                      for (int j = 0; j < 16; ++j)
                        if (col + j < p.N) pre[j] = static_cast<uint16_t>(pack2(gp[j], 0.f, p.fmt) & 0xFFFFu);
                    }
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
                }
              }
              if (p.residual != nullptr && row_ok) {
                if (full_slice) {
#pragma unroll
                  for (int j4 = 0; j4 < 2; ++j4) {
                    const uint32_t w[4] = {r4[j4].x, r4[j4].y, r4[j4].z, r4[j4].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      const float2 x = unpack2(w[q], p.fmt);
                      if (ACT == 2) {
                        f[j4 * 8 + q * 2] *= gelu_erf_grad(x.x);
                        f[j4 * 8 + q * 2 + 1] *= gelu_erf_grad(x.y);
                      } else if (ACT == 3) {
                        f[j4 * 8 + q * 2] *= x.x;
                        f[j4 * 8 + q * 2 + 1] *= x.y;
                      } else {
                        f[j4 * 8 + q * 2] += x.x;
                        f[j4 * 8 + q * 2 + 1] += x.y;
                      }
                    }
                  }
                } else {
                  const uint16_t* r16 = static_cast<const uint16_t*>(p.residual) + grow * p.ldr + col;
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (col + j < p.N) {
                      const float x = unpack2(static_cast<uint32_t>(r16[j]), p.fmt).x;
                      if (ACT == 2) f[j] *= gelu_erf_grad(x);
                      else if (ACT == 3) f[j] *= x;
                      else f[j] += x;
                    }
                }
              }
              // stage the slice into its half of the 32 x 32 box (swizzled exactly as the output tensor map expects)
              if (OUT_F32) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<float4*>(buf + lane * 128 + (((hs * 4 + j) ^ (lane & 7)) << 4)) =
                      make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  uint4 w;
                  w.x = pack2(f[8 * j], f[8 * j + 1], p.fmt);
                  w.y = pack2(f[8 * j + 2], f[8 * j + 3], p.fmt);
                  w.z = pack2(f[8 * j + 4], f[8 * j + 5], p.fmt);
                  w.w = pack2(f[8 * j + 6], f[8 * j + 7], p.fmt);
                  *reinterpret_cast<uint4*>(buf + lane * 64 + (((hs * 2 + j) ^ ((lane >> 1) & 3)) << 4)) = w;
                }
              }
            }
            if (hs == 1) {   // box complete (columns past N, if any, are clipped by the TMA store)
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                const int bcol = tile_col + bx * 32;
                if (RED) ptx::tma_reduce_add_2d(&tmap_out, buf, bcol, m_tile * kBM + quarter * 32);
                else ptx::tma_store_2d(&tmap_out, buf, bcol, m_tile * kBM + quarter * 32);
                ptx::bulk_commit_group();
              }
              ++nstore;
            }
          }
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (lane == 0) ptx::bulk_wait_group_read<0>();  // the staging buffers must outlive the TMA reads
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (CTAS == 2) ptx::cluster_sync_all();  // neither CTA may retire (or free TMEM) while the pair's MMAs can still touch it
  if (warp == 1) {
    ptx::tc_fence_after();
    if (CTAS == 2) ptx::tmem_dealloc_2cta(tmem_base, 512);
    else ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ldot
