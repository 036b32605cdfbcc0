// Fused Linear + residual + LayerNorm, CTA-PAIR form for N = 768 (see linear_ln.cuh for the arithmetic and the epilogue).
//
// linear_ln_kernel's single-CTA 128 x 256 tiles pull 48 KB of operands per 64-wide K block (96 B/clk/SM) and are bound by
// that traffic, not by their epilogue.  Here a cluster is 2 x 3 CTAs: three PAIRS (cta_group::2), one
// per 256-column tile, each pair owning 256 rows.  A CTA stages its own 128 rows of A and HALF of its pair's weight tile
// (32 KB per K block instead of 48), the pair's leader issues 256 x 256 x 16 MMAs, and the per-row LayerNorm statistics
// travel between the three CTAs that hold the same rows (cluster ranks mh, 2 + mh, 4 + mh) exactly as before.  The
// residual still rides through the tensor core (256 x 64 x 64 products with I_64, 32 identity rows per CTA).
// Clusters of 6 CTAs: 22 are resident on a B200 (132 of 148 SMs; clusters of 3: 135).
#pragma once
#include "linear_ln.cuh"

namespace ldot {

struct Ln2Smem {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (kLinBN / 2) * kBK * 2;   // this CTA's half of the n-tile's 256 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingPerWarp = 2048;         // one 32 x 32 16-bit box
  static constexpr int kStages = 6;
  static constexpr int kStagingOffset = kStages * kStageBytes;
  static constexpr int kStatsOffset = kStagingOffset + kLinEpiWarps * kStagingPerWarp;
  // [2 tile parities][C * 2 column halves][128 rows] float2
  static constexpr int kStatsBytes = 2 * kLnMaxCluster * 2 * kBM * 8;
  static constexpr int kVecOffset = kStatsOffset + kStatsBytes;   // bias | gamma | beta of this CTA's 256 columns
  static constexpr int kVecBytes = 3 * kLinBN * 4;
  static constexpr int kBarOffset = kVecOffset + kVecBytes;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 4 + 2) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
};
static_assert(Ln2Smem::kDynamic <= 227 * 1024, "linear+LN pair kernel shared memory");

__global__ void __launch_bounds__(kLinThreads, 1)
linear_ln2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_eye,
                 const __grid_constant__ CUtensorMap tmap_out, const LnSched sched, const LnParams p) {
  using SM = Ln2Smem;
  constexpr int kStages = SM::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * SM::kABytes;
  float2* stats = reinterpret_cast<float2*>(smem + SM::kStatsOffset);
  float* vec = reinterpret_cast<float*>(smem + SM::kVecOffset);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* sbar = tempty + 2;  // [2] statistics of tile parity b complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sbar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int C = 3;                                           // n-tiles (N = 768)
  const int crank = static_cast<int>(ptx::cluster_ctarank());    // 0 .. 5
  const int rank = crank >> 1;                                   // n-tile of this CTA's pair
  const int mh = crank & 1;                                      // row half of the pair's 256 rows; 0 = the pair's leader
  const uint32_t leader = static_cast<uint32_t>(crank & ~1);
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << leader);
  const int cluster_id = blockIdx.x / (2 * C);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], kLinEpiWarps * 2);   // (leader's copy: both CTAs' epilogue warps arrive)
      // C > 1: armed per tile with expect_tx of the C * 8 * 32 partials (8 B each) it will receive through st.async;
      // C == 1 (N <= 256, not a cluster launch): every epilogue lane stores locally and arrives
      ptx::mbar_init(&sbar[i], 1u);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_res);
    ptx::prefetch_tmap(&tmap_eye);
    ptx::prefetch_tmap(&tmap_out);
  }
  if (warp == 1) {
    ptx::tmem_alloc_2cta(tmem_ptr, 512);
    ptx::tmem_relinquish_2cta();
  }
  // the per-column vectors of this CTA's n-tile never change: stage them once (columns past N read as 0 / unused)
  for (int i = threadIdx.x; i < kLinBN; i += blockDim.x) {
    const int col = rank * kLinBN + i;
    const bool ok = col < p.N;
    vec[i] = ok && p.bias != nullptr ? __ldg(p.bias + col) : 0.f;
    vec[kLinBN + i] = ok ? __ldg(p.gamma + col) : 0.f;
    vec[2 * kLinBN + i] = ok ? __ldg(p.beta + col) : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // remote arrives must find initialised barriers
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA: its rows, its half of W)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int m_pair = cluster_id; m_pair < sched.m_tiles; m_pair += sched.num_clusters) {
        const int a_row = (m_pair * 2 + mh) * kBM;
        for (int kb = 0; kb < sched.k_blocks + sched.res_blocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t lbar = ptx::mapa(ptx::smem_u32(&full[stage]), leader);   // bytes of both CTAs: leader's barrier
          uint8_t* sa = smem_a + stage * SM::kABytes;
          uint8_t* sb = smem_b + stage * SM::kBBytes;
          if (kb < sched.k_blocks) {
            if (mh == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * SM::kStageBytes);
            ptx::tma_load_2d_2cta(sa, &tmap_a, lbar, kb * kBK, a_row, ptx::kEvictNormal);
            ptx::tma_load_2d_2cta(sb, &tmap_w, lbar, kb * kBK, rank * kLinBN + mh * (kLinBN / 2), ptx::kEvictLast);
          } else {  // residual columns [64 j, 64 j + 64) of this n-tile against I_64 (each CTA stages 32 of its rows)
            const int j = kb - sched.k_blocks;
            if (mh == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * (SM::kABytes + (kBK / 2) * kBK * 2));
            ptx::tma_load_2d_2cta(sa, &tmap_res, lbar, rank * kLinBN + j * kBK, a_row, ptx::kEvictNormal);
            ptx::tma_load_2d_2cta(sb, &tmap_eye, lbar, 0, mh * (kBK / 2), ptx::kEvictLast);
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
    // ------------------------------------------------------------------ MMA issuer (pair leaders only)
    if (mh == 0 && ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int m_pair = cluster_id; m_pair < sched.m_tiles; m_pair += sched.num_clusters) {
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kLinBN);
        for (int kb = 0; kb < sched.k_blocks + sched.res_blocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a + stage * SM::kABytes));
          const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b + stage * SM::kBBytes));
          if (kb < sched.k_blocks) {
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              ptx::mma_f16_ss_2cta(tmem_d, adesc + 2 * k, bdesc + 2 * k, sched.idesc, (kb | k) != 0 ? 1u : 0u);
          } else {
            // residual block j: a 256 x 64 x 64 product with I_64 into accumulator columns [64 j, 64 j + 64) of both CTAs
            const uint32_t tmem_dj = tmem_d + static_cast<uint32_t>((kb - sched.k_blocks) * kBK);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              ptx::mma_f16_ss_2cta(tmem_dj, adesc + 2 * k, bdesc + 2 * k, sched.idesc_res, 1u);
          }
          ptx::mma_commit_2cta(&empty[stage], pair_mask);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::mma_commit_2cta(&tfull[as], pair_mask);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps, two passes per tile)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    uint8_t* staging = smem + SM::kStagingOffset + (warp - 2) * SM::kStagingPerWarp;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int col_base = rank * kLinBN + half * (kLinBN / 2);
    int nchunks = (p.N - col_base + 31) / 32;  // N % 32 == 0 (host-checked): chunks are full or absent
    nchunks = nchunks < 0 ? 0 : (nchunks > 4 ? 4 : nchunks);
    const float inv_n = 1.0f / static_cast<float>(p.N);
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t tempty_leader[2] = {ptx::mapa(ptx::smem_u32(&tempty[0]), leader), ptx::mapa(ptx::smem_u32(&tempty[1]), leader)};
    for (int m_pair = cluster_id; m_pair < sched.m_tiles; m_pair += sched.num_clusters) {
      const int m_tile = m_pair * 2 + mh;   // this CTA's 128-row block
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = lane_base + static_cast<uint32_t>(as * kLinBN + half * (kLinBN / 2));

      // x = acc (+ residual, already accumulated by the MMA) + bias for one 32-column chunk (both passes).  TMEM loads are
      // software-pipelined in both passes: chunk cc + 1 is in flight while chunk cc is consumed.
      uint32_t v2[2][32];
      auto add_bias = [&](int cc, const uint32_t (&v)[32], float (&f)[32]) {
        const float4* b4 = reinterpret_cast<const float4*>(vec + half * (kLinBN / 2) + cc * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j];  // (broadcast LDS.128)
          f[4 * j] = __uint_as_float(v[4 * j]) + b.x;
          f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
          f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
          f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
        }
      };

      // ---- pass 1: partial row statistics, broadcast to the whole cluster
      float s1 = 0.f, s2 = 0.f;
      if (nchunks > 0) ptx::tmem_ld32(taddr, v2[0]);
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc >= nchunks) break;
        ptx::tmem_ld_wait();
        // (after the last chunk of pass 1 the first chunk of pass 2 is fetched: it overlaps the statistics exchange)
        ptx::tmem_ld32(taddr + (cc + 1 < nchunks ? cc + 1 : 0) * 32, v2[(cc + 1) & 1]);
        float f[32];
        add_bias(cc, v2[cc & 1], f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s1 += f[j];
          s2 = fmaf(f[j], f[j], s2);
        }
      }
      const int first2 = nchunks & 1;   // v2 buffer that holds pass 2's chunk 0
      {
        const uint32_t slot = ptx::smem_u32(stats + (as * 2 * kLnMaxCluster + rank * 2 + half) * kBM + row);
        const uint32_t bar = ptx::smem_u32(&sbar[as]);
        // st.async: every partial carries its own complete_tx to the destination CTA's barrier - no release fence on
        // this side, no cluster-scope acquire polling on the other
        // the three CTAs that hold the other column tiles of THESE rows: cluster ranks mh, 2 + mh, 4 + mh
        if (warp == 2 && lane == 0) ptx::mbar_arrive_expect_tx(&sbar[as], static_cast<uint32_t>(C * kLinEpiWarps * 32 * 8));
        for (int c = 0; c < C; ++c)
          ptx::st_async_f32x2(ptx::mapa(slot, static_cast<uint32_t>(2 * c + mh)), s1, s2,
                              ptx::mapa(bar, static_cast<uint32_t>(2 * c + mh)));
      }
      ptx::mbar_wait(&sbar[as], aphase);
      float t1 = 0.f, t2 = 0.f;
      for (int i = 0; i < 2 * C; ++i) {
        const float2 s = stats[(as * 2 * kLnMaxCluster + i) * kBM + row];
        t1 += s.x;
        t2 += s.y;
      }
      const float mean = t1 * inv_n;
      const float var = fmaxf(t2 * inv_n - mean * mean, 0.f);
      const float rstd = rsqrtf(var + kLnEpsF);

      // ---- pass 2: normalise, scale, pack, store
      if (nchunks == 0) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(tempty_leader[as]);
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc >= nchunks) break;
        const int col = col_base + cc * 32;
        ptx::tmem_ld_wait();
        if (cc + 1 < nchunks) {
          if ((first2 + cc + 1) & 1) ptx::tmem_ld32(taddr + (cc + 1) * 32, v2[1]);
          else ptx::tmem_ld32(taddr + (cc + 1) * 32, v2[0]);
        } else {  // accumulator drained for good
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(tempty_leader[as]);
        }
        float f[32];
        if ((first2 + cc) & 1) add_bias(cc, v2[1], f);
        else add_bias(cc, v2[0], f);
        const float4* g4 = reinterpret_cast<const float4*>(vec + kLinBN + half * (kLinBN / 2) + cc * 32);
        const float4* e4 = reinterpret_cast<const float4*>(vec + 2 * kLinBN + half * (kLinBN / 2) + cc * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 g = g4[j];
          const float4 b = e4[j];
          f[4 * j] = fmaf((f[4 * j] - mean) * rstd, g.x, b.x);
          f[4 * j + 1] = fmaf((f[4 * j + 1] - mean) * rstd, g.y, b.y);
          f[4 * j + 2] = fmaf((f[4 * j + 2] - mean) * rstd, g.z, b.z);
          f[4 * j + 3] = fmaf((f[4 * j + 3] - mean) * rstd, g.w, b.w);
        }
        if (lane == 0) ptx::bulk_wait_group_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = pack2(f[8 * j], f[8 * j + 1], p.fmt);
          w.y = pack2(f[8 * j + 2], f[8 * j + 3], p.fmt);
          w.z = pack2(f[8 * j + 4], f[8 * j + 5], p.fmt);
          w.w = pack2(f[8 * j + 6], f[8 * j + 7], p.fmt);
          *reinterpret_cast<uint4*>(staging + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = w;
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tmap_out, staging, col, m_tile * kBM + quarter * 32);
          ptx::bulk_commit_group();
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
    if (lane == 0) ptx::bulk_wait_group_read<0>();
    __syncwarp();
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // no CTA may retire while a peer can still write its statistics table / barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace ldot
