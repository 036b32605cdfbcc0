// A-stationary variant of the search main loop: the 128-query tile lives in TENSOR MEMORY for a whole unit.
//
// Why: with both operands streamed through shared memory (gemm_tc.cuh) every 128 x 256 score tile pulls
// (128 + 256) x K x 2 bytes out of L2 - 85 FLOP per L2 byte, i.e. ~10.6 TB/s of L2->SM traffic at the rate the
// tensor pipe can go: the L2 fabric, not the MMA, sets the pace.  Queries are reused against every index row, so
// they are parked where the MMA can read them for free: TMEM.  A [128 x K] 16-bit tile is 128 lanes x K/2 columns
// (K = 768 -> 384 of the 512 columns); the remaining 128 columns hold two 64-column fp32 accumulators.  Only index
// rows stream (TMA -> smem ring -> tcgen05.mma A-from-TMEM): 128 FLOP per L2 byte, and the whole 192 KB of shared
// memory is pipeline depth for the one stream that matters.
//
// Roles: warp 0 TMA producer (index rows), warp 1 MMA issuer, warps 2..5 epilogue.  At the start of every unit the
// epilogue warps copy their query rows global -> registers -> TMEM (tcgen05.st; a warp can only touch its own lane
// quarter) and release the MMA warp through `a_ready`.
#pragma once
#include "gemm_tc.cuh"

namespace ldot {

constexpr int kTsBN = 64;          // index rows per MMA (UMMA N)
constexpr int kTsKbPerStage = 4;   // 64-wide K blocks per smem stage (4 x 8 KB)
constexpr int kTsStages = 6;       // 6 x 32 KB = 192 KB in flight
constexpr int kTsFastKb = 12;      // K blocks of the compile-time specialised kernel (d = 768): two tiles per ring revolution
constexpr int kTsACols = 384;      // TMEM columns holding the query tile (K <= 768)
constexpr int kTsMaxD = kTsACols * 2;

struct TsQueries {
  const uint16_t* q16;  // [nq, d] 16-bit queries, row-major
  int nq, d;
};

struct TsSmem {
  static constexpr int kKbBytes = kTsBN * kBK * 2;  // 8 KB
  static constexpr int kStageBytes = kTsKbPerStage * kKbBytes;
  static constexpr int kBarOffset = kTsStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * kTsStages + 5) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
};

// TMA issue for one index tile on the compile-time path (see ts_issue_tile): one 3-D TMA instruction per stage lands
// kTsKbPerStage consecutive [64 rows x 64 K] swizzled tiles.
template <int KB, int H>
__device__ __forceinline__ void ts_load_tile(uint32_t ph, uint8_t* smem, const CUtensorMap* tmap3, uint64_t* full,
                                             uint64_t* empty, int row0) {
  constexpr int kStagesPerTile = KB / kTsKbPerStage;
#pragma unroll
  for (int s = 0; s < kStagesPerTile; ++s) {
    const int stage = H * kStagesPerTile + s;
    ptx::mbar_wait(&empty[stage], ph ^ 1);
    ptx::mbar_arrive_expect_tx(&full[stage], TsSmem::kStageBytes);
    ptx::tma_load_3d(smem + stage * TsSmem::kStageBytes, tmap3, &full[stage], 0, row0, s * kTsKbPerStage,
                     ptx::kEvictNormal);
  }
}

// MMA issue for one index tile when the vector length is known at compile time (KB 64-wide K blocks, a multiple of
// kTsKbPerStage with 2 * KB / kTsKbPerStage == kTsStages: a tile occupies exactly half of the ring, H selects which
// half and which accumulator).  Every
// shared-memory descriptor, TMEM column and barrier address is the per-CTA base plus an immediate, so the single
// issuing thread spends ~3 instructions per tcgen05.mma - with 64-row tiles an MMA lasts only 32 cycles and the
// generic loop below (runtime stage / descriptor arithmetic, ~15 instructions per MMA) is issue-bound at half rate.
template <int KB, int H>
__device__ __forceinline__ void ts_issue_tile(uint32_t ph, uint64_t desc0, uint32_t tmem_a, uint32_t tmem_d0,
                                              uint64_t* full, uint64_t* empty, uint64_t* tfull, uint64_t* tempty,
                                              uint32_t idesc) {
  constexpr int kStagesPerTile = KB / kTsKbPerStage;
  ptx::mbar_wait(&tempty[H], ph ^ 1);
  ptx::tc_fence_after();
  const uint32_t tmem_d = tmem_d0 + static_cast<uint32_t>(H * kTsBN);
#pragma unroll
  for (int s = 0; s < kStagesPerTile; ++s) {
    constexpr int kDescPerKb = TsSmem::kKbBytes >> 4;
    const int stage = H * kStagesPerTile + s;
    ptx::mbar_wait(&full[stage], ph);
    ptx::tc_fence_after();
#pragma unroll
    for (int j = 0; j < kTsKbPerStage; ++j) {
#pragma unroll
      for (int k = 0; k < kBK / kUmmaK; ++k) {
        const int kb = s * kTsKbPerStage + j;
        ptx::mma_f16_ts(tmem_d, tmem_a + static_cast<uint32_t>(kb * (kBK / 2) + k * (kUmmaK / 2)),
                        desc0 + static_cast<uint64_t>((stage * kTsKbPerStage + j) * kDescPerKb + 2 * k), idesc,
                        (kb | k) != 0 ? 1u : 0u);
      }
    }
    ptx::mma_commit(&empty[stage]);
  }
  ptx::mma_commit(&tfull[H]);
}

template <class Epi, int KB>
__global__ void __launch_bounds__(kGemmThreads, 1)
coarse_ts_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_b3,
                 const GemmSched sched, const TsQueries tq, const typename Epi::Params ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TsSmem::kBarOffset);
  uint64_t* empty = full + kTsStages;
  uint64_t* tfull = empty + kTsStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_ready = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kb2_count = (sched.k_blocks + kTsKbPerStage - 1) / kTsKbPerStage;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTsStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], 128);
    }
    ptx::mbar_init(a_ready, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) ptx::prefetch_tmap(KB > 0 ? &tmap_b3 : &tmap_b);
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_acc = tmem_base + kTsACols;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (index rows only)
    if constexpr (KB > 0) {
      if (ptx::elect_one()) {
        uint32_t tile_idx = 0;
        for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
          const UnitInfo u = unit_info(sched, unit);
          for (int it = 0; it < u.n_tile_end - u.n_tile_begin; ++it, ++tile_idx) {
            const int row0 = unit_tile(u, it) * sched.tile_stride * kTsBN;
            const uint32_t ph = (tile_idx >> 1) & 1u;
            if ((tile_idx & 1u) == 0) ts_load_tile<KB, 0>(ph, smem, &tmap_b3, full, empty, row0);
            else ts_load_tile<KB, 1>(ph, smem, &tmap_b3, full, empty, row0);
          }
        }
      }
    } else if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
        const UnitInfo u = unit_info(sched, unit);
        for (int it = 0; it < u.n_tile_end - u.n_tile_begin; ++it) {
          const int nt = unit_tile(u, it);
          for (int kb2 = 0; kb2 < kb2_count; ++kb2) {
            const int kb0 = kb2 * kTsKbPerStage;
            const int nkb = sched.k_blocks - kb0 < kTsKbPerStage ? sched.k_blocks - kb0 : kTsKbPerStage;
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            ptx::mbar_arrive_expect_tx(&full[stage], nkb * TsSmem::kKbBytes);
            for (int j = 0; j < nkb; ++j)
              ptx::tma_load_2d(smem + stage * TsSmem::kStageBytes + j * TsSmem::kKbBytes, &tmap_b, &full[stage],
                               (kb0 + j) * kBK, nt * sched.tile_stride * kTsBN, ptx::kEvictNormal);
            if (++stage == kTsStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (A from TMEM, B from smem)
    if constexpr (KB > 0) {
      static_assert(KB % kTsKbPerStage == 0 && 2 * (KB / kTsKbPerStage) == kTsStages, "a tile must fill half the ring");
      if (ptx::elect_one()) {
        const uint64_t desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
        uint32_t tile_idx = 0, uphase = 0;
        for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
          const UnitInfo u = unit_info(sched, unit);
          ptx::mbar_wait(a_ready, uphase);  // this unit's query tile is in TMEM
          uphase ^= 1;
          ptx::tc_fence_after();
          for (int it = u.n_tile_begin; it < u.n_tile_end; ++it, ++tile_idx) {
            const uint32_t ph = (tile_idx >> 1) & 1u;
            if ((tile_idx & 1u) == 0)
              ts_issue_tile<KB, 0>(ph, desc0, tmem_base, tmem_acc, full, empty, tfull, tempty, sched.idesc);
            else
              ts_issue_tile<KB, 1>(ph, desc0, tmem_base, tmem_acc, full, empty, tfull, tempty, sched.idesc);
          }
        }
      }
    } else if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0, uphase = 0;
      for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
        const UnitInfo u = unit_info(sched, unit);
        ptx::mbar_wait(a_ready, uphase);  // this unit's query tile is in TMEM
        uphase ^= 1;
        ptx::tc_fence_after();
        for (int nt = u.n_tile_begin; nt < u.n_tile_end; ++nt) {
          ptx::mbar_wait(&tempty[as], aphase ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_acc + static_cast<uint32_t>(as * kTsBN);
          for (int kb2 = 0; kb2 < kb2_count; ++kb2) {
            const int kb0 = kb2 * kTsKbPerStage;
            const int nkb = sched.k_blocks - kb0 < kTsKbPerStage ? sched.k_blocks - kb0 : kTsKbPerStage;
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            for (int j = 0; j < nkb; ++j) {
              const uint64_t bdesc =
                  ptx::make_smem_desc_sw128(ptx::smem_u32(smem + stage * TsSmem::kStageBytes + j * TsSmem::kKbBytes));
              const uint32_t a_col = tmem_base + static_cast<uint32_t>((kb0 + j) * (kBK / 2));
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k)
                ptx::mma_f16_ts(tmem_d, a_col + k * (kUmmaK / 2), bdesc + 2 * k, sched.idesc,
                                ((kb0 + j) | k) != 0 ? 1u : 0u);
            }
            ptx::mma_commit(&empty[stage]);
            if (++stage == kTsStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          ptx::mma_commit(&tfull[as]);
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (+ query tile loader)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int as = 0;
    uint32_t aphase = 0;
    typename Epi::State st;
    for (int unit = blockIdx.x; unit < sched.num_units; unit += gridDim.x) {
      const UnitInfo u = unit_info(sched, unit);
      // All MMAs of the previous unit have completed (its last tfull was consumed below), so the query tile can be
      // replaced.  Thread = query row: 64 K-elements (128 B) per tcgen05.st of 32 columns.
      {
        const int q = u.m_tile * kBM + row;
        const uint4* src = reinterpret_cast<const uint4*>(tq.q16 + static_cast<size_t>(q) * tq.d);
        for (int kb = 0; kb < sched.k_blocks; ++kb) {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 w = make_uint4(0, 0, 0, 0);
            if (q < tq.nq && kb * kBK + j * 8 < tq.d) w = __ldg(src + kb * 8 + j);  // d % 8 == 0
            v[4 * j] = w.x;
            v[4 * j + 1] = w.y;
            v[4 * j + 2] = w.z;
            v[4 * j + 3] = w.w;
          }
          ptx::tmem_st32(lane_base + static_cast<uint32_t>(kb * (kBK / 2)), v);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(a_ready);
      }
      Epi::unit_begin(st, ep, u, row);
      for (int it = 0; it < u.n_tile_end - u.n_tile_begin; ++it) {
        const int nt = unit_tile(u, it);
        ptx::mbar_wait(&tfull[as], aphase);
        ptx::tc_fence_after();
        Epi::tile(st, ep, u, row, nt * sched.tile_stride, lane_base + kTsACols + static_cast<uint32_t>(as * kTsBN));
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
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ldot
