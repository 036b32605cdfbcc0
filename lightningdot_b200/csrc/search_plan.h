// Host-side plan of one flat inner-product search call: tile / chunk decomposition and workspace layout.
#pragma once
#include <cstddef>
#include <cstdint>

namespace ldot {

constexpr int kSearchBN = 256;     // index rows per MMA tile (UMMA N) of the streamed kernel (d > 768)
constexpr int kSearchStages = 4;   // smem ring depth: 4 x (16 KB queries + 32 KB index rows)

struct SearchPlan {
  int kprime;          // coarse candidates kept per query (>= k)
  int epl;             // list entries per lane in the compaction (cap = epl * 32 >= 2 * kprime)
  int cap;
  int kp_pad;          // kprime rounded up to a power of two (bitonic sort width)
  int a_in_tmem;       // 1: A-stationary kernel (queries in tensor memory, 64-row index tiles); 0: streamed 128 x 256
  int bn;              // index rows per MMA tile
  int m_tiles, n_tiles, tiles_per_unit, chunks, num_units;
  int groups;          // candidate-list groups of the select stage (> 1: two-level select)
  // threshold pre-pass over a strided sample of the index tiles (0 = off, small indexes)
  int sample;          // 1: run it
  int s_stride;        // every s_stride-th index tile is sampled
  int s_tiles, s_tiles_per_unit, s_chunks, s_units;
  int s_kprime;        // the s_kprime-th best sample score of a query seeds its threshold
  size_t off_tmax;     // tile maxima of the pre-pass [m_tiles * 128][s_tiles] fp32
  size_t off_q16, off_qstats, off_qmu, off_gtau, off_flagcnt, off_cnt, off_sel_idx, off_sel_cmin, off_sel_key, off_l2_ent,
      off_l2_cnt, off_cand;
  size_t total_bytes;
};

struct SearchArgs {
  const float* q;       // [nq, d] fp32 queries (device)
  long long nq;
  const float* x;       // [n, d] fp32 master index (device)
  const void* x16;      // [n, d] centred 16-bit copy (device)
  const float* mu;      // [d] centring vector
  const float* xstats;  // [2] residual / norm maxima from index_prepare
  long long n;
  int d, k, coarse_k, coarse_dtype;
  long long id_offset;
  float* out_scores;    // [nq, k]
  long long* out_idx;   // [nq, k]
  int* out_flags;       // [nq]
  int* out_flag_count;  // [1] (device) or null
  void* ws;
  size_t ws_bytes;
  void* stream;
  // sharded search in two phases over the SAME workspace (0 = the whole search in one call):
  //   1  query prepare .. select, then bound_out[q] = what bound_m rows of this shard are guaranteed to reach
  //   2  rescore + rank + certificate, pruned by tau[q] (a lower bound of the global k-th best score; may be null)
  int phase = 0;
  int bound_m = 0;
  float* bound_out = nullptr;
  const float* tau = nullptr;
};

int search_make_plan(SearchPlan* pl, long long nq, long long n, int d, int k, int coarse_k, int sms);
int search_run(const SearchArgs& a);
size_t index_prepare_workspace_bytes(int d);
int index_prepare_run(const float* x, long long n, int d, int coarse_dtype, int center, void* x16, float* mu,
                      float* xstats, void* ws, size_t ws_bytes, void* stream);
size_t exact_workspace_bytes(long long n);
int exact_run(const float* q, long long nf, const float* x, long long n, int d, int k, long long id_offset,
              float* out_scores, long long* out_idx, void* ws, size_t ws_bytes, void* stream);
int merge_run(const float* scores, const long long* idx, int W, long long nq, int k, long long ws_s, long long ws_i,
              float* out_scores,
              long long* out_idx, void* stream);

}  // namespace ldot
