// BertSelfAttention in ONE kernel (uniter_model/model/layer.py:60-101):
//
//     ctx[tokens, H] = softmax( (x Wq^T + bq)(x Wk^T + bk)^T / 8 + (1 - mask) * -10000 ) (x Wv^T + bv)     per head
//
// The fused Q|K|V projection (a [tokens, 768] x [2304, 768]^T tcgen05 GEMM) used to write its [tokens, 2304] result to
// HBM for a separate attention kernel to read back: 1.47 GB out + 1.47 GB in per layer at 10 000 captions, and a kernel
// that is instruction-bound on its own.  Here a CTA pair (cta_group::2) owns a 256-row x 192-column output tile whose
// columns are Q_h | K_h | V_h of ONE head - the three 64-row slices of the stacked [3H, H] weight are fetched by three TMA
// boxes per stage, no weight permutation - so one 128-row accumulator holds everything attention needs for the
// floor(128 / S) whole sequences that start in it.  The epilogue warps
//   (A) drain the accumulator: + bias, round to 16 bit, into three padded shared-memory tiles (Q, K, V: 160 x 64),
//       and hand the TMEM buffer straight back to the MMA warp (the next tile's main loop runs under phase B);
//   (B) per (sequence, 16-query-row block): S = Q K^T on mma.sync m16n8k16 (ldmatrix from the tiles), scale + additive
//       mask, fp32 softmax, P V, normalise, and write the 16 x 64 context block to ctx.
// Q, K, V never reach HBM.  M tiles are laid out in whole sequences: tile t starts at token t * R with
// R = floor(128 / S) * S useful rows (TMA fetches 128 rows from there; the tail rows belong to the next tile and are
// ignored), so any S <= 128 works: S = 32 -> 4 sequences per tile, no waste; S = 37 (UNITER image: [CLS] + 36 regions) ->
// 3 sequences, 13 % of the tensor work unused.
// Warp roles and the operand pipeline are those of linear_tc.cuh's CTA-pair form.
#pragma once
#include "linear_tc.cuh"
#include "mma_sync.cuh"
#include "rowops.cuh"

namespace ldot {

constexpr int kQaBN = 192;            // Q_h | K_h | V_h
constexpr int kQaStages = 5;
constexpr int kQaEpiWarps = 8;
constexpr int kQaThreads = 32 * (2 + kQaEpiWarps);
constexpr int kQaTileRows = 160;      // 128 accumulator rows + 32 zero rows (a key block may run past row 127)

struct QaSched {
  int m_tiles;        // 128-row tiles of whole sequences
  int heads;
  int num_tiles;      // ceil(m_tiles / 2) * heads pair-tiles, head fastest
  int k_blocks;
  int rows_per_tile;  // R = seq_per_tile * S
  int seq_per_tile;
  uint32_t idesc;     // 256 x 192 x 16, cta_group::2
  int yield_mma;      // how the MMA issuer makes room for the epilogue warps' mma.sync bursts (see qa_hold): 0 not at all
                      // (1.27 ms per 10k-caption layer), 1 waits mid-tile for the previous tile's attention (1.13),
                      // 2 = default: skips a K block while a burst is announced (1.09)
};

struct QaParams {
  const float* bias;        // [3 H]: bq | bk | bv
  const long long* mask;    // [B, S] attention_mask (1 = attend)
  uint16_t* ctx;            // [B * S, H] 16-bit
  int B, S, H;
};

struct QaSmem {
  static constexpr int kABytes = kBM * kBK * 2;               // 16 KB
  static constexpr int kBBytes = (kQaBN / 2) * kBK * 2;       // 12 KB: this CTA's 96 weight rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTileOffset = kQaStages * kStageBytes;
  static constexpr int kMatBytes = kQaTileRows * kRowPad * 2;  // one of Q / K / V
  static constexpr int kMaskOffset = kTileOffset + 3 * kMatBytes;       // additive mask of the tile's 128 token rows (fp32)
  static constexpr int kBarOffset = kMaskOffset + kBM * 4;
  static constexpr int kTotal = kBarOffset + (2 * kQaStages + 4 + 2) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
};
static_assert(QaSmem::kDynamic <= 227 * 1024, "qkv+attention kernel shared memory");
static_assert(QaSmem::kBBytes % 1024 == 0 && QaSmem::kTileOffset % 1024 == 0, "swizzle atoms must stay 1024 B aligned");

// Legacy mma.sync shares the tensor pipe with tcgen05.mma and is starved while the pair's main loop streams (measured:
// the first HMMA of a burst waits ~10x longer than the rest; tensor pipe 54 % active).  Epilogue warps therefore announce
// their HMMA bursts in a counter that lives in the LEADER CTA's shared memory; the MMA issuer does not queue the next
// K block while the counter is non-zero.  `hold` = shared::cluster address of the counter, 0 = protocol off.
__device__ __forceinline__ void qa_hold(uint32_t hold, int delta, int lane) {
  if (hold == 0) return;
  __syncwarp();
  if (lane == 0) asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(hold), "r"(delta) : "memory");
}

// One 16-query-row block of one (sequence, head): sQ / sK / sV point at the sequence's first row inside the tiles.
template <int SPAD, int FMT>
__device__ __forceinline__ void qa_attention_block(uint16_t* sQ, const uint16_t* sK, const uint16_t* sV,
                                                   const float* madd, uint16_t* __restrict__ out,
                                                   int H, int S, int qrow0, int lane, uint32_t hold) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t qa[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int r = qrow0 + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int c = ks * 16 + 8 * (lane >> 4);
    ldsm_x4(qa[ks], ptx::smem_u32(sQ + r * kRowPad + c));
  }
  constexpr int NT = SPAD / 8;
  float sc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
  qa_hold(hold, 1, lane);
#pragma unroll
  for (int n2 = 0; n2 < NT / 2; ++n2) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t kb[4];
      const int r = n2 * 16 + (lane & 7) + 8 * (lane >> 4);
      const int c = ks * 16 + 8 * ((lane >> 3) & 1);
      ldsm_x4(kb, ptx::smem_u32(sK + r * kRowPad + c));
      mma16816<FMT>(sc[2 * n2], qa[ks], kb[0], kb[1]);
      mma16816<FMT>(sc[2 * n2 + 1], qa[ks], kb[2], kb[3]);
    }
  }
  qa_hold(hold, -1, lane);
  // scale by 1 / sqrt(64), additive mask (uniter_model/model/model.py:362-365), softmax over the keys
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = n * 8 + t * 2 + j;
      const float add = col < S ? madd[col] : -INFINITY;   // (0 / -10000, staged by the drain)
      sc[n][j] = fmaf(sc[n][j], 0.125f, add);
      sc[n][2 + j] = fmaf(sc[n][2 + j], 0.125f, add);
      mx0 = fmaxf(mx0, sc[n][j]);
      mx1 = fmaxf(mx1, sc[n][2 + j]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xFFFFFFFFu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xFFFFFFFFu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      sc[n][j] = __expf(sc[n][j] - mx0);
      sc[n][2 + j] = __expf(sc[n][2 + j] - mx1);
      sum0 += sc[n][j];
      sum1 += sc[n][2 + j];
    }
  }
  sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, 1);
  sum0 += __shfl_xor_sync(0xFFFFFFFFu, sum0, 2);
  sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, 1);
  sum1 += __shfl_xor_sync(0xFFFFFFFFu, sum1, 2);
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  qa_hold(hold, 1, lane);
#pragma unroll
  for (int kk = 0; kk < SPAD / 16; ++kk) {
    uint32_t pa[4];
    pa[0] = pk2<FMT>(sc[2 * kk][0], sc[2 * kk][1]);
    pa[1] = pk2<FMT>(sc[2 * kk][2], sc[2 * kk][3]);
    pa[2] = pk2<FMT>(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
    pa[3] = pk2<FMT>(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
    for (int d2 = 0; d2 < 4; ++d2) {
      uint32_t vb[4];
      const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int c = d2 * 16 + 8 * (lane >> 4);
      ldsm_x4_t(vb, ptx::smem_u32(sV + r * kRowPad + c));
      mma16816<FMT>(o[2 * d2], pa, vb[0], vb[1]);
      mma16816<FMT>(o[2 * d2 + 1], pa, vb[2], vb[3]);
    }
  }
  qa_hold(hold, -1, lane);
  const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
  // stage the block's 16 x 64 output in its own (now dead) Q rows - only rows of THIS sequence: a block that runs past the
  // sequence's end overlaps the next sequence's Q rows, which another warp still reads - then 16-byte coalesced stores
  __syncwarp();
  const int r0 = qrow0 + g, r1 = qrow0 + g + 8;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (r0 < S) *reinterpret_cast<uint32_t*>(sQ + r0 * kRowPad + n * 8 + t * 2) = pk2<FMT>(o[n][0] * inv0, o[n][1] * inv0);
    if (r1 < S) *reinterpret_cast<uint32_t*>(sQ + r1 * kRowPad + n * 8 + t * 2) = pk2<FMT>(o[n][2] * inv1, o[n][3] * inv1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    const int r = qrow0 + (idx >> 3), c = (idx & 7) * 8;
    if (r < S)
      *reinterpret_cast<uint4*>(out + static_cast<long long>(r) * H + c) = *reinterpret_cast<const uint4*>(sQ + r * kRowPad + c);
  }
}

template <int SPAD, int FMT>
__global__ void __launch_bounds__(kQaThreads, 1)
qkv_attn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const QaSched sched,
                const QaParams p) {
  using SM = QaSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kQaStages * SM::kABytes;
  uint16_t* tiles = reinterpret_cast<uint16_t*>(smem + SM::kTileOffset);   // Q | K | V, kQaTileRows x kRowPad each
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::kBarOffset);
  uint64_t* empty = full + kQaStages;
  uint64_t* tfull = empty + kQaStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* adone = tempty + 2;   // [2] (leader's copy is used): attention phase of accumulator parity a finished, both CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(adone + 2);
  uint32_t* hold_ctr = tmem_ptr + 1;
  float* maskadd = reinterpret_cast<float*>(smem + SM::kMaskOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(ptx::cluster_ctarank());   // 0 = leader of the pair
  const int first_tile = static_cast<int>(blockIdx.x) / 2;
  const int tile_step = static_cast<int>(gridDim.x) / 2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kQaStages; ++i) {
      ptx::mbar_init(&full[i], 1);
      ptx::mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull[i], 1);
      ptx::mbar_init(&tempty[i], kQaEpiWarps * 2);
      ptx::mbar_init(&adone[i], kQaEpiWarps * 2);
    }
    *hold_ctr = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    ptx::tmem_alloc_2cta(tmem_ptr, 512);
    ptx::tmem_relinquish_2cta();
  }
  // rows 128 .. 159 of the three tiles are never written by the drain: keep them zero (finite keys / values)
  for (int i = threadIdx.x; i < 3 * (kQaTileRows - kBM) * kRowPad / 2; i += blockDim.x) {
    const int mat = i / ((kQaTileRows - kBM) * kRowPad / 2), rem = i - mat * ((kQaTileRows - kBM) * kRowPad / 2);
    reinterpret_cast<uint32_t*>(tiles + mat * kQaTileRows * kRowPad + kBM * kRowPad)[rem] = 0u;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs of the pair)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first_tile; t < sched.num_tiles; t += tile_step) {
        const int m_pair = t / sched.heads, head = t - m_pair * sched.heads;
        const int a_row = (m_pair * 2 + rank) * sched.rows_per_tile;
        for (int kb = 0; kb < sched.k_blocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * SM::kABytes;
          uint8_t* sb = smem_b + stage * SM::kBBytes;
          const uint32_t lbar = ptx::mapa(ptx::smem_u32(&full[stage]), 0);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * SM::kStageBytes);
          ptx::tma_load_2d_2cta(sa, &tmap_a, lbar, kb * kBK, a_row, ptx::kEvictNormal);
          // this CTA's 96 rows of the [Q_h | K_h | V_h] weight tile: tile rows [96 rank, 96 rank + 96) in 32-row boxes
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int j0 = rank * (kQaBN / 2) + i * 32;
            const int w_row = (j0 >> 6) * p.H + head * kHeadDim + (j0 & 63);
            ptx::tma_load_2d_2cta(sb + i * 32 * kBK * 2, &tmap_w, lbar, kb * kBK, w_row, ptx::kEvictLast);
          }
          if (++stage == kQaStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      uint32_t dphase[2] = {0, 0};
      int it_count = 0;
      for (int t = first_tile; t < sched.num_tiles; t += tile_step, ++it_count) {
        ptx::mbar_wait(&tempty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kLinBN);
        for (int kb = 0; kb < sched.k_blocks; ++kb) {
          if (sched.yield_mma == 2) {
            // (checked per K block; per instruction measured worse: 1.13 -> 1.62 ms, the issuer then trails every warp's burst)
            while (*reinterpret_cast<volatile uint32_t*>(hold_ctr) != 0) {
            }
          }
          if (sched.yield_mma == 1 && kb == sched.k_blocks / 2 && it_count >= 1) {
            // legacy mma.sync is starved while tcgen05.mma streams: let the epilogue warps' attention of the PREVIOUS tile
            // (accumulator parity as ^ 1) finish before the second half of this main loop is queued
            ptx::mbar_wait(&adone[as ^ 1], dphase[as ^ 1]);
            dphase[as ^ 1] ^= 1;
          }
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint64_t adesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a + stage * SM::kABytes));
          const uint64_t bdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b + stage * SM::kBBytes));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            ptx::mma_f16_ss_2cta(tmem_d, adesc + 2 * k, bdesc + 2 * k, sched.idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::mma_commit_2cta(&empty[stage], 0x3);
          if (++stage == kQaStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::mma_commit_2cta(&tfull[as], 0x3);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps): drain, then attention
    const int ew = warp - 2;
    const int quarter = warp & 3;           // TMEM lane quarter this warp may read
    const int half = ew >> 2;               // accumulator columns [96 half, 96 half + 96)
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t tempty_addr[2] = {ptx::mapa(ptx::smem_u32(&tempty[0]), 0), ptx::mapa(ptx::smem_u32(&tempty[1]), 0)};
    const uint32_t adone_addr[2] = {ptx::mapa(ptx::smem_u32(&adone[0]), 0), ptx::mapa(ptx::smem_u32(&adone[1]), 0)};
    const long long T = static_cast<long long>(p.B) * p.S;
    const uint32_t hold = sched.yield_mma == 2 ? ptx::mapa(ptx::smem_u32(hold_ctr), 0) : 0u;
    const int qblocks = (p.S + 15) / 16;
    const int items = sched.seq_per_tile * qblocks;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = first_tile; t < sched.num_tiles; t += tile_step) {
      const int m_pair = t / sched.heads, head = t - m_pair * sched.heads;
      const int m_tile = m_pair * 2 + rank;
      ptx::mbar_wait(&tfull[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = lane_base + static_cast<uint32_t>(as * kLinBN + half * (kQaBN / 2));
      // additive mask of the tile's token rows (attention_mask is [B, S] contiguous = one value per token)
      long long mval = 1;
      if (half == 0) {
        const long long tok = static_cast<long long>(m_tile) * sched.rows_per_tile + row;
        if (row < sched.rows_per_tile && tok < T) mval = __ldg(p.mask + tok);
      }
      // ---- (A) accumulator -> + bias -> 16 bit -> Q / K / V tiles
      uint32_t v2[2][32];
      ptx::tmem_ld32(taddr, v2[0]);
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) {
        const int c0 = half * (kQaBN / 2) + cc * 32;      // first tile column of this chunk
        const int mat = c0 >> 6, within = c0 & 63;
        const float* b = p.bias + mat * p.H + head * kHeadDim + within;
        float4 b4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(b) + j);
        ptx::tmem_ld_wait();
        if (cc + 1 < 3) {
          ptx::tmem_ld32(taddr + (cc + 1) * 32, v2[(cc + 1) & 1]);
        } else {   // drained: the MMA warp may start the tile after next in this buffer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_remote(tempty_addr[as]);
        }
        const uint32_t (&v)[32] = v2[cc & 1];
        uint16_t* dst = tiles + mat * kQaTileRows * kRowPad + row * kRowPad + within;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 w;
          w.x = pk2<FMT>(__uint_as_float(v[8 * j]) + b4[2 * j].x, __uint_as_float(v[8 * j + 1]) + b4[2 * j].y);
          w.y = pk2<FMT>(__uint_as_float(v[8 * j + 2]) + b4[2 * j].z, __uint_as_float(v[8 * j + 3]) + b4[2 * j].w);
          w.z = pk2<FMT>(__uint_as_float(v[8 * j + 4]) + b4[2 * j + 1].x, __uint_as_float(v[8 * j + 5]) + b4[2 * j + 1].y);
          w.w = pk2<FMT>(__uint_as_float(v[8 * j + 6]) + b4[2 * j + 1].z, __uint_as_float(v[8 * j + 7]) + b4[2 * j + 1].w);
          *reinterpret_cast<uint4*>(dst + 8 * j) = w;
        }
      }
      if (half == 0) maskadd[row] = mval != 0 ? 0.f : -10000.f;
      asm volatile("bar.sync 1, %0;" ::"n"(kQaEpiWarps * 32) : "memory");   // tiles complete
      // ---- (B) attention: (sequence, 16-row query block) items over the 8 warps
      if (m_tile < sched.m_tiles) {
        for (int it = ew; it < items; it += kQaEpiWarps) {
          const int s = it / qblocks, qb = it - s * qblocks;
          const long long seq = static_cast<long long>(m_tile) * sched.seq_per_tile + s;
          if (seq >= p.B) break;
          const int r0 = s * p.S;
          uint16_t* sQ = tiles + r0 * kRowPad;
          qa_attention_block<SPAD, FMT>(sQ, sQ + kQaTileRows * kRowPad, sQ + 2 * kQaTileRows * kRowPad, maskadd + r0,
                                        p.ctx + seq * p.S * p.H + head * kHeadDim, p.H, p.S, qb * 16, lane, hold);
        }
      }
      if (sched.yield_mma == 1) {
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_remote(adone_addr[as]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kQaEpiWarps * 32) : "memory");   // tiles free for the next drain
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace ldot
