// Warp-specialised persistent tcgen05 GEMM main loop for sm_100a.
//
//   C[M, N] = A[M, K] * B[N, K]^T      A, B: 16-bit (bf16 / fp16), K-major (row-major [rows, K]); fp32 accumulate
//
// Both the index search (A = queries, B = index rows) and every encoder Linear (A = activations, B = nn.Linear
// weight [out, in]) have exactly this shape, so one main loop serves the whole hot path.  What differs is the
// epilogue, supplied as a policy class `Epi`:
//
//   struct Epi {
//     struct Params { ... };                       // kernel argument (POD)
//     struct State  { ... };                       // per-thread registers that live across the tiles of a unit
//     static __device__ void unit_begin(State&, const Params&, const UnitInfo&, int row);   // row: 0..127 in tile
//     static __device__ void tile(State&, const Params&, const UnitInfo&, int row, int n_tile, uint32_t taddr);
//     static __device__ void unit_end(State&, const Params&, const UnitInfo&, int row);
//   };
//
// `taddr` is the TMEM address of this warp's 32-lane slice of the 128 x BN fp32 accumulator: the thread reads
// ITS row (lane = row) with tcgen05.ld 32x32b, i.e. one thread owns one output row - one query in the search
// epilogue, one token in the encoder epilogues.
//
// Roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer (one elected
// lane), warps 2..5 = epilogue (TMEM lane quarter = warp % 4).  Pipelines: STAGES-deep smem ring (full/empty
// mbarriers, TMA -> MMA) and a 2-deep TMEM accumulator ring (tfull/tempty, MMA -> epilogue) so the epilogue of
// tile i overlaps the MMAs of tile i+1.
//
// Work decomposition: a *unit* is one 128-row M tile times a run of consecutive N tiles ("chunk").  Units are
// numbered m-fastest (unit = chunk * m_tiles + m_tile) and dealt round-robin to the persistent CTAs, so CTAs that
// run at the same time read the same B chunk (L2 reuse of the index / the weights) with different A tiles.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace ldot {

constexpr int kBM = 128;       // rows of A per tile (UMMA M)
constexpr int kBK = 64;        // K elements per smem stage: 64 x 2 B = one 128 B swizzle row
constexpr int kUmmaK = 16;     // K per tcgen05.mma for 16-bit inputs
constexpr int kGemmThreads = 192;

struct GemmSched {
  int m_tiles;         // ceil(M / 128)
  int n_tiles;         // ceil(N / BN)
  int tiles_per_unit;  // N tiles per unit (chunk length)
  int chunks;          // ceil(n_tiles / tiles_per_unit)
  int num_units;       // m_tiles * chunks
  int k_blocks;        // ceil(K / 64)
  int tile_stride;     // N tile t of the schedule is tile t * tile_stride of B (> 1: strided sample of the index)
  uint32_t idesc;      // tcgen05 instruction descriptor (dtype, M=128, N=BN)
};

struct UnitInfo {
  int unit, m_tile, chunk, n_tile_begin, n_tile_end;
  int rot;  // the unit visits its tiles starting at n_tile_begin + rot (wrapping): CTAs that share a B chunk in L2
            // then pull different lines at any one time instead of hammering the same L2 slice together
};

// i-th tile visited by a unit
__device__ __forceinline__ int unit_tile(const UnitInfo& u, int i) {
  const int len = u.n_tile_end - u.n_tile_begin;
  int t = i + u.rot;
  if (t >= len) t -= len;
  return u.n_tile_begin + t;
}

__device__ __forceinline__ UnitInfo unit_info(const GemmSched& s, int unit) {
  UnitInfo u;
  u.unit = unit;
  u.chunk = unit / s.m_tiles;
  u.m_tile = unit - u.chunk * s.m_tiles;
  u.n_tile_begin = u.chunk * s.tiles_per_unit;
  int e = u.n_tile_begin + s.tiles_per_unit;
  u.n_tile_end = e < s.n_tiles ? e : s.n_tiles;
  const int len = u.n_tile_end - u.n_tile_begin;
  u.rot = len > 0 ? static_cast<int>((static_cast<unsigned>(u.m_tile) * 29u) % static_cast<unsigned>(len)) : 0;
  return u;
}

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kBM * kBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBK * 2;   // 32 KB at BN = 256
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  // full[STAGES], empty[STAGES], tfull[2], tempty[2], tmem ptr
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for manual 1024 B alignment
};

template <class Epi, int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmSched sched, const typename Epi::Params ep) {
  using SM = GemmSmem<BN, STAGES>;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * SM::kABytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
        const UnitInfo u = unit_info(sched, unit);
        for (int it = 0; it < u.n_tile_end - u.n_tile_begin; ++it) {
          const int nt = unit_tile(u, it);
          for (int kb = 0; kb < sched.k_blocks; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full[stage], SM::kStageBytes);
            ptx::tma_load_2d(smem_a + stage * SM::kABytes, &tmap_a, &full[stage], kb * kBK, u.m_tile * kBM,
                             ptx::kEvictLast);
            ptx::tma_load_2d(smem_b + stage * SM::kBBytes, &tmap_b, &full[stage], kb * kBK,
                             nt * sched.tile_stride * BN, ptx::kEvictNormal);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
        const UnitInfo u = unit_info(sched, unit);
        for (int nt = u.n_tile_begin; nt < u.n_tile_end; ++nt) {
          ptx::mbar_wait(&tempty[as], aphase ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
          for (int kb = 0; kb < sched.k_blocks; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint64_t adesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a + stage * SM::kABytes));
            const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b + stage * SM::kBBytes));
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              // advance 16 K-elements = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
              ptx::mma_f16_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, sched.idesc, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::mma_commit(&empty[stage]);  // frees the smem slot when these MMAs have read it
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          ptx::mma_commit(&tfull[as]);  // accumulator complete
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 rows)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int as = 0;
    uint32_t aphase = 0;
    typename Epi::State st;
    for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
      const UnitInfo u = unit_info(sched, unit);
      Epi::unit_begin(st, ep, u, row);
      for (int it = 0; it < u.n_tile_end - u.n_tile_begin; ++it) {
        const int nt = unit_tile(u, it);
        ptx::mbar_wait(&tfull[as], aphase);
        ptx::tc_fence_after();
        Epi::tile(st, ep, u, row, nt * sched.tile_stride, lane_base + static_cast<uint32_t>(as * BN));
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tempty[as]);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      Epi::unit_end(st, ep, u, row);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace ldot
