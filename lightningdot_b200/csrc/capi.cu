// extern "C" surface of libldot_sm100a.so (declared in include/ldot.h) + shared host helpers.
#include "../../include/ldot.h"
#include <cstring>
#include <mutex>
#include <vector>
#include "host_common.h"
#include "prof.h"
#include "search_plan.h"

namespace ldot {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_kmajor_16b(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                         uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(kErrCuda, "cuTensorMapEncodeTiled entry point not available");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  LDOT_REQUIRE(row_stride_bytes % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  LDOT_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  // the element type only matters for OOB fill / arithmetic; bf16 and fp16 are both plain 2-byte moves
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(kErrCuda, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return kOk;
}

int make_tmap_kblocks_16b(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                          uint32_t box_rows, uint32_t box_kb) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(kErrCuda, "cuTensorMapEncodeTiled entry point not available");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && row_stride_bytes % 16 == 0, "TMA alignment");
  LDOT_REQUIRE(cols % 64 == 0 && box_rows >= 1 && box_rows <= 256 && box_kb >= 1 && box_kb <= 256, "bad 3-D box");
  const cuuint64_t gdim[3] = {64, rows, cols / 64};
  const cuuint64_t gstride[2] = {row_stride_bytes, 128};
  const cuuint32_t box[3] = {64, box_rows, box_kb};
  const cuuint32_t estride[3] = {1, 1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(kErrCuda, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
  return kOk;
}

int make_tmap_store(CUtensorMap* out, const void* base, uint32_t elt_bytes, uint64_t rows, uint64_t cols,
                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, bool as_float) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(kErrCuda, "cuTensorMapEncodeTiled entry point not available");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && row_stride_bytes % 16 == 0,
               "TMA store target must be 16-byte aligned with a 16-byte multiple row pitch");
  const uint32_t inner = box_cols * elt_bytes;
  LDOT_REQUIRE((elt_bytes == 2 || elt_bytes == 4) && (inner == 64 || inner == 128), "unsupported TMA store box");
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  const CUtensorMapDataType dt = elt_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                 : as_float     ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  const CUresult r = fn(out, dt, 2,
                        const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        inner == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(kErrCuda, "cuTensorMapEncodeTiled (store) failed with CUresult %d", (int)r);
  return kOk;
}

// ---------------------------------------------------------------------------------------------- launch accounting
namespace {
struct ProfRecord {
  int cls;
  cudaEvent_t e0, e1;
};
struct ProfState {
  std::mutex mu;
  bool enabled = false;
  long long launches[kKcCount] = {0};
  double flops[kKcCount] = {0};
  double bytes[kKcCount] = {0};
  double ms[kKcCount] = {0};            // folded in by prof_collect
  long long timed[kKcCount] = {0};      // launches that carry an event pair
  std::vector<ProfRecord> open;         // recorded, not yet folded
  std::vector<cudaEvent_t> pool;        // reusable events
};
ProfState& prof_state() {
  static ProfState s;
  return s;
}
}  // namespace

const char* kernel_class_name(int c) {
  static const char* names[kKcCount] = {"coarse_score_topk", "select", "rescore", "query_prepare", "index_prepare",
                                        "exact_scan", "merge", "linear_tcgen05", "attention", "layernorm", "embed",
                                        "cast", "nll", "optim", "qkv_attention"};
  return c >= 0 && c < kKcCount ? names[c] : "?";
}

void prof_begin(int cls, cudaStream_t st, double flops, double bytes, void** token) {
  ProfState& S = prof_state();
  std::lock_guard<std::mutex> lock(S.mu);
  S.launches[cls] += 1;
  *token = nullptr;
  if (!S.enabled) return;
  S.flops[cls] += flops;   // work totals cover exactly the timed launches
  S.bytes[cls] += bytes;
  ProfRecord r;
  r.cls = cls;
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) {
    if (!S.pool.empty()) {
      ev[i] = S.pool.back();
      S.pool.pop_back();
    } else if (cudaEventCreate(&ev[i]) != cudaSuccess) {
      return;
    }
  }
  r.e0 = ev[0];
  r.e1 = ev[1];
  cudaEventRecord(r.e0, st);
  S.open.push_back(r);
  *token = reinterpret_cast<void*>(static_cast<uintptr_t>(S.open.size()));  // index + 1
}

void prof_end(void* token, cudaStream_t st) {
  ProfState& S = prof_state();
  std::lock_guard<std::mutex> lock(S.mu);
  const size_t i = static_cast<size_t>(reinterpret_cast<uintptr_t>(token)) - 1;
  if (i < S.open.size()) cudaEventRecord(S.open[i].e1, st);
}

static int prof_collect() {
  ProfState& S = prof_state();
  std::lock_guard<std::mutex> lock(S.mu);
  for (const ProfRecord& r : S.open) {
    LDOT_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    LDOT_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    S.ms[r.cls] += ms;
    S.timed[r.cls] += 1;
    S.pool.push_back(r.e0);
    S.pool.push_back(r.e1);
  }
  S.open.clear();
  return kOk;
}

int device_sm_count(int* out) {
  static int cached[64] = {0};
  int dev = 0;
  LDOT_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && cached[dev]) {
    *out = cached[dev];
    return kOk;
  }
  int sms = 0;
  LDOT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (dev < 64) cached[dev] = sms;
  *out = sms;
  return kOk;
}

}  // namespace ldot

#include "dropout.cuh"
namespace ldot {
int linear_run(const void* a, long long lda, const void* w, long long ldw, const float* bias, const void* residual,
               long long ldr, void* out, long long ldo, long long M, int N, int K, int fmt, int act, int out_f32,
               void* stream, const DropKey* drop = nullptr, void* pre = nullptr, long long ld_pre = 0);
}

namespace ldot {
int linear_ln_run(const void* a, long long lda, const void* w, long long ldw, const float* bias, const void* residual,
                  long long ldr, const float* gamma, const float* beta, void* out, long long ldo, long long M, int N,
                  int K, int fmt, void* stream);
}  // namespace ldot
#include "encoder_params.h"
namespace ldot {
int qkv_attention_run(const void* x, long long ldx, const void* w, long long ldw, const float* bias, const long long* mask,
                      void* ctx, int B, int S, int H, int heads, int K, int fmt, void* stream);
int attention_run(const void* qkv, const long long* mask, void* ctx, int B, int S, int H, int heads, int q_rows, int fmt,
                  void* stream, float drop_p = 0.f, unsigned long long seed = 0, int site = 0);
int layernorm_run(const void* in, long long ld_in, int in_f32, const float* gamma, const float* beta, void* out,
                  long long ld_out, long long rows, int H, int fmt, void* stream);
int embed_text_run(const long long* ids, const long long* pos_ids, long long pos_batch_stride, const void* word,
                   const void* pos, const void* type0, const float* gamma, const float* beta, void* out, int B, int L,
                   int out_seq, int H, int vocab, int max_pos, int fmt, void* stream);
int embed_image_run(const EmbedImageParams& p, int H, int fmt, void* stream);
int cast_run(const float* in, void* out, long long n, int fmt, void* stream);
int split16_run(const float* in, long long rows, int K, int side, void* out, void* stream);
int nll_run(const float* s1, const float* s2, float w, const long long* pos, long long bq, long long bc, int reduction,
            float* s_out, float* row_loss, int* row_correct, float* loss, long long* correct, void* stream);
}

#include "train_params.h"

using namespace ldot;

extern "C" {

int ldot_abi_version(void) { return LDOT_ABI_VERSION; }

const char* ldot_last_error(void) { return error_buffer(); }

int ldot_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  LDOT_CUDA(cudaGetDevice(&dev));
  LDOT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  LDOT_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return set_error(kErrArch, "libldot_sm100a needs an sm_100 device, found sm_%d%d", major, minor);
  return kOk;
}

int ldot_prof_num_classes(void) { return kKcCount; }

const char* ldot_prof_class_name(int32_t cls) { return kernel_class_name(cls); }

int ldot_prof_enable(int32_t on) {
  if (int e = prof_collect()) return e;
  ProfState& S = prof_state();
  std::lock_guard<std::mutex> lock(S.mu);
  S.enabled = on != 0;
  return kOk;
}

int ldot_prof_reset(void) {
  if (int e = prof_collect()) return e;
  ProfState& S = prof_state();
  std::lock_guard<std::mutex> lock(S.mu);
  for (int c = 0; c < kKcCount; ++c) {
    S.launches[c] = S.timed[c] = 0;
    S.flops[c] = S.bytes[c] = S.ms[c] = 0.0;
  }
  return kOk;
}

int ldot_prof_read(int32_t n_classes, int64_t* launches, int64_t* timed_launches, double* ms, double* flops,
                   double* bytes) {
  LDOT_REQUIRE(n_classes >= 0 && launches && timed_launches && ms && flops && bytes, "null pointer argument");
  if (int e = prof_collect()) return e;
  ProfState& S = prof_state();
  std::lock_guard<std::mutex> lock(S.mu);
  for (int c = 0; c < n_classes && c < kKcCount; ++c) {
    launches[c] = S.launches[c];
    timed_launches[c] = S.timed[c];
    ms[c] = S.ms[c];
    flops[c] = S.flops[c];
    bytes[c] = S.bytes[c];
  }
  return kOk;
}

size_t ldot_index_prepare_workspace_bytes(int64_t n, int32_t d) {
  (void)n;
  return index_prepare_workspace_bytes(d);
}

int ldot_index_prepare(const float* d_x, int64_t n, int32_t d, int32_t coarse_dtype, int32_t center, void* d_x16,
                       float* d_mu, float* d_xstats, void* d_ws, size_t ws_bytes, void* stream) {
  LDOT_REQUIRE(d_x && d_x16 && d_mu && d_xstats && d_ws, "null pointer argument");
  return index_prepare_run(d_x, n, d, coarse_dtype, center, d_x16, d_mu, d_xstats, d_ws, ws_bytes, stream);
}

size_t ldot_flatip_search_workspace_bytes(int64_t nq, int64_t n, int32_t d, int32_t k, int32_t coarse_k) {
  SearchPlan pl;
  int sms = 0;
  if (device_sm_count(&sms) != kOk) sms = 148;
  if (search_make_plan(&pl, nq, n, d, k, coarse_k, sms) != kOk) return 0;
  return pl.total_bytes;
}

int ldot_flatip_search(const float* d_q, int64_t nq, const float* d_x, const void* d_x16, const float* d_mu,
                       const float* d_xstats, int64_t n, int32_t d, int32_t k, int32_t coarse_k,
                       int32_t coarse_dtype, int64_t id_offset, float* d_out_scores, int64_t* d_out_idx,
                       int32_t* d_out_flags, int32_t* d_out_flag_count, void* d_ws, size_t ws_bytes, void* stream) {
  return ldot_flatip_search_phase(d_q, nq, d_x, d_x16, d_mu, d_xstats, n, d, k, coarse_k, coarse_dtype, id_offset, d_out_scores,
                                  d_out_idx, d_out_flags, d_out_flag_count, d_ws, ws_bytes, 0, 0, nullptr, nullptr, stream);
}

int ldot_flatip_search_phase(const float* d_q, int64_t nq, const float* d_x, const void* d_x16, const float* d_mu,
                             const float* d_xstats, int64_t n, int32_t d, int32_t k, int32_t coarse_k,
                             int32_t coarse_dtype, int64_t id_offset, float* d_out_scores, int64_t* d_out_idx,
                             int32_t* d_out_flags, int32_t* d_out_flag_count, void* d_ws, size_t ws_bytes, int32_t phase,
                             int32_t bound_m, float* d_bound, const float* d_tau, void* stream) {
  LDOT_REQUIRE(d_q && d_x && d_x16 && d_mu && d_xstats && d_out_scores && d_out_idx && d_out_flags && d_ws,
               "null pointer argument");
  LDOT_REQUIRE(phase >= 0 && phase <= 2, "phase must be 0 (whole search), 1 (through select + bound) or 2 (rescore)");
  static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
  SearchArgs a;
  a.q = d_q;
  a.nq = nq;
  a.x = d_x;
  a.x16 = d_x16;
  a.mu = d_mu;
  a.xstats = d_xstats;
  a.n = n;
  a.d = d;
  a.k = k;
  a.coarse_k = coarse_k;
  a.coarse_dtype = coarse_dtype;
  a.id_offset = id_offset;
  a.out_scores = d_out_scores;
  a.out_idx = reinterpret_cast<long long*>(d_out_idx);
  a.out_flags = d_out_flags;
  a.out_flag_count = d_out_flag_count;
  a.ws = d_ws;
  a.ws_bytes = ws_bytes;
  a.stream = stream;
  a.phase = phase;
  a.bound_m = bound_m;
  a.bound_out = d_bound;
  a.tau = d_tau;
  return search_run(a);
}

size_t ldot_flatip_exact_workspace_bytes(int64_t n) { return exact_workspace_bytes(n); }

int ldot_flatip_exact(const float* d_q, int64_t nq, const float* d_x, int64_t n, int32_t d, int32_t k,
                      int64_t id_offset, float* d_out_scores, int64_t* d_out_idx, void* d_ws, size_t ws_bytes,
                      void* stream) {
  LDOT_REQUIRE(d_q && d_x && d_out_scores && d_out_idx && d_ws, "null pointer argument");
  return exact_run(d_q, nq, d_x, n, d, k, id_offset, d_out_scores, reinterpret_cast<long long*>(d_out_idx), d_ws,
                   ws_bytes, stream);
}

int ldot_topk_merge(const float* d_scores, const int64_t* d_idx, int32_t world, int64_t nq, int32_t k,
                    int64_t shard_stride_scores, int64_t shard_stride_idx, float* d_out_scores, int64_t* d_out_idx,
                    void* stream) {
  LDOT_REQUIRE(d_scores && d_idx && d_out_scores && d_out_idx, "null pointer argument");
  return merge_run(d_scores, reinterpret_cast<const long long*>(d_idx), world, nq, k, shard_stride_scores,
                   shard_stride_idx, d_out_scores,
                   reinterpret_cast<long long*>(d_out_idx), stream);
}

int ldot_linear(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias,
                const void* d_residual, int64_t ldr, void* d_out, int64_t ldo, int64_t M, int32_t N, int32_t K,
                int32_t dtype, int32_t act, int32_t out_f32, void* stream) {
  LDOT_REQUIRE(d_a && d_w && d_out, "null pointer argument");
  return linear_run(d_a, lda, d_w, ldw, d_bias, d_residual, ldr, d_out, ldo, M, N, K, dtype, act, out_f32, stream);
}

int ldot_linear_ln(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias, const void* d_residual,
                   int64_t ldr, const float* d_gamma, const float* d_beta, void* d_out, int64_t ldo, int64_t M, int32_t N,
                   int32_t K, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_a && d_w && d_out && d_gamma && d_beta, "null pointer argument");
  return linear_ln_run(d_a, lda, d_w, ldw, d_bias, d_residual, ldr, d_gamma, d_beta, d_out, ldo, M, N, K, dtype, stream);
}

int ldot_layernorm(const void* d_in, int64_t ld_in, int32_t in_f32, const float* d_gamma, const float* d_beta,
                   void* d_out, int64_t ld_out, int64_t rows, int32_t H, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_in && d_gamma && d_beta && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return layernorm_run(d_in, ld_in, in_f32, d_gamma, d_beta, d_out, ld_out, rows, H, dtype, stream);
}

int ldot_embed_text(const int64_t* d_ids, const int64_t* d_pos_ids, int64_t pos_batch_stride, const void* d_word,
                    const void* d_pos, const void* d_type0, const float* d_gamma, const float* d_beta, void* d_out,
                    int32_t B, int32_t L, int32_t out_seq, int32_t H, int32_t vocab, int32_t max_pos, int32_t dtype,
                    void* stream) {
  LDOT_REQUIRE(d_ids && d_pos_ids && d_word && d_pos && d_type0 && d_gamma && d_beta && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return embed_text_run(reinterpret_cast<const long long*>(d_ids), reinterpret_cast<const long long*>(d_pos_ids),
                        pos_batch_stride, d_word, d_pos, d_type0, d_gamma, d_beta, d_out, B, L, out_seq, H, vocab,
                        max_pos, dtype, stream);
}

int ldot_embed_image(const float* d_lin, const float* d_box, const float* d_img_g, const float* d_img_b,
                     const float* d_pos_w, const float* d_pos_bias, const float* d_pos_g, const float* d_pos_b,
                     const float* d_type1, const float* d_ln_g, const float* d_ln_b, void* d_out, int32_t B, int32_t R,
                     int32_t out_seq, int32_t row_offset, int32_t H, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_lin && d_box && d_img_g && d_img_b && d_pos_w && d_pos_bias && d_pos_g && d_pos_b && d_type1 &&
                   d_ln_g && d_ln_b && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  EmbedImageParams p;
  p.lin = d_lin; p.box = d_box; p.img_g = d_img_g; p.img_b = d_img_b; p.pos_w = d_pos_w; p.pos_bias = d_pos_bias;
  p.pos_g = d_pos_g; p.pos_b = d_pos_b; p.type1 = d_type1; p.ln_g = d_ln_g; p.ln_b = d_ln_b;
  p.out = static_cast<uint16_t*>(d_out); p.B = B; p.R = R; p.out_seq = out_seq; p.row_offset = row_offset;
  return embed_image_run(p, H, dtype, stream);
}

int ldot_attention(const void* d_qkv, const int64_t* d_mask, void* d_ctx, int32_t B, int32_t S, int32_t H,
                   int32_t heads, int32_t q_rows, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_qkv && d_mask && d_ctx, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return attention_run(d_qkv, reinterpret_cast<const long long*>(d_mask), d_ctx, B, S, H, heads, q_rows, dtype, stream);
}

int ldot_qkv_attention(const void* d_x, int64_t ldx, const void* d_w, int64_t ldw, const float* d_bias, const int64_t* d_mask,
                       void* d_ctx, int32_t B, int32_t S, int32_t H, int32_t heads, int32_t K, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_x && d_w && d_bias && d_mask && d_ctx, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return qkv_attention_run(d_x, ldx, d_w, ldw, d_bias, reinterpret_cast<const long long*>(d_mask), d_ctx, B, S, H, heads, K,
                           dtype, stream);
}

int ldot_attention_train(const void* d_qkv, const int64_t* d_mask, void* d_ctx, int32_t B, int32_t S, int32_t H,
                         int32_t heads, float drop_p, uint64_t seed, int32_t site, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_qkv && d_mask && d_ctx, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return attention_run(d_qkv, reinterpret_cast<const long long*>(d_mask), d_ctx, B, S, H, heads, S, dtype, stream, drop_p,
                       seed, site);
}

int ldot_dropout(const void* d_x, const void* d_res, void* d_out, int64_t rows, int32_t cols, int64_t ld, float p,
                 uint64_t seed, int32_t site, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_x && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return dropout_run(d_x, d_res, d_out, rows, cols, ld, p, seed, site, dtype, stream);
}

int ldot_cast_f32(const float* d_in, void* d_out, int64_t n, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_in && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return cast_run(d_in, d_out, n, dtype, stream);
}

int ldot_split16(const float* d_in, int64_t rows, int32_t K, int32_t side, void* d_out, void* stream) {
  LDOT_REQUIRE(d_in && d_out, "null pointer argument");
  return split16_run(d_in, rows, K, side, d_out, stream);
}

int ldot_inbatch_nll(const float* d_scores, const float* d_scores_cap, float cap_weight, const int64_t* d_pos,
                     int64_t bq, int64_t bc, int32_t reduction, float* d_scores_out, float* d_row_loss,
                     int32_t* d_row_correct, float* d_loss, int64_t* d_correct, void* stream) {
  LDOT_REQUIRE(d_scores && d_pos && d_scores_out && d_row_loss && d_row_correct && d_loss && d_correct,
               "null pointer argument");
  return nll_run(d_scores, d_scores_cap, cap_weight, reinterpret_cast<const long long*>(d_pos), bq, bc, reduction,
                 d_scores_out, d_row_loss, d_row_correct, d_loss, reinterpret_cast<long long*>(d_correct), stream);
}

// ---- training step ----------------------------------------------------------------------
int ldot_gemm(const void* d_a, int64_t lda, int32_t a_mn, const void* d_b, int64_t ldb, int32_t b_mn,
              const float* d_bias, const void* d_aux, int64_t ld_aux, void* d_out, int64_t ldo, int64_t M, int32_t N,
              int64_t K, int32_t dtype, int32_t epi, int32_t out_f32, int32_t accumulate, void* stream) {
  LDOT_REQUIRE(d_a && d_b && d_out, "null pointer argument");
  return gemm_run(d_a, lda, a_mn, d_b, ldb, b_mn, d_bias, d_aux, ld_aux, d_out, ldo, M, N, K, dtype, epi, out_f32,
                  accumulate, stream);
}

int ldot_layernorm_bwd(const void* d_dy, int64_t ld_dy, int32_t dy_f32, const void* d_x, int64_t ld_x, int32_t x_f32,
                       const float* d_gamma, void* d_dx, int64_t ld_dx, int32_t dx_f32, float* d_dgamma,
                       float* d_dbeta, float* d_dxsum, int64_t rows, int32_t H, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_dy && d_x && d_gamma && d_dx && d_dgamma && d_dbeta, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  LnBwdParams p;
  p.dy = d_dy; p.ld_dy = ld_dy; p.dy_f32 = dy_f32; p.x = d_x; p.ld_x = ld_x; p.x_f32 = x_f32; p.gamma = d_gamma;
  p.dx = d_dx; p.ld_dx = ld_dx; p.dx_f32 = dx_f32; p.dgamma = d_dgamma; p.dbeta = d_dbeta; p.dxsum = d_dxsum;
  p.rows = rows; p.fmt = dtype; p.dx_masked = nullptr; p.drop = make_drop_key(0.f, 0, 0);
  return ln_bwd_run(p, H, stream);
}

int ldot_attention_bwd(const void* d_qkv, const int64_t* d_mask, const void* d_ctx, const void* d_dctx, void* d_dqkv,
                       int32_t B, int32_t S, int32_t H, int32_t heads, float drop_p, uint64_t seed, int32_t site,
                       int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_qkv && d_mask && d_ctx && d_dctx && d_dqkv, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return attention_bwd_run(d_qkv, reinterpret_cast<const long long*>(d_mask), d_ctx, d_dctx, d_dqkv, B, S, H, heads,
                           dtype, stream, drop_p, seed, site);
}

int ldot_gelu(const void* d_x, void* d_out, int64_t n, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_x && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return gelu_run(d_x, nullptr, d_out, n, 0, dtype, stream);
}

int ldot_gelu_grad(void* d_z_gp, void* d_out, int64_t n, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_z_gp && d_out && d_z_gp != d_out, "gelu_grad: two distinct buffers are required");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return gelu_run(d_z_gp, d_z_gp, d_out, n, 2, dtype, stream);
}

int ldot_gelu_bwd(const void* d_x, const void* d_dy, void* d_dx, int64_t n, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_x && d_dy && d_dx, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return gelu_run(d_x, d_dy, d_dx, n, 1, dtype, stream);
}

int ldot_colsum16(const void* d_in, int64_t ld, int64_t rows, int32_t N, float* d_out, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_in && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return colsum16_run(d_in, ld, rows, N, d_out, dtype, stream);
}

int ldot_embed_text_sum(const int64_t* d_ids, const int64_t* d_pos_ids, int64_t pos_batch_stride, const void* d_word,
                        const void* d_pos, const void* d_type0, float* d_out, int32_t B, int32_t L, int32_t H,
                        int32_t vocab, int32_t max_pos, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_ids && d_pos_ids && d_word && d_pos && d_type0 && d_out, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return embed_text_sum_run(reinterpret_cast<const long long*>(d_ids), reinterpret_cast<const long long*>(d_pos_ids),
                            pos_batch_stride, d_word, d_pos, d_type0, d_out, B, L, H, vocab, max_pos, dtype, stream);
}

int ldot_embed_scatter(const float* d_dx, const int64_t* d_ids, const int64_t* d_pos_ids, int64_t pos_batch_stride,
                       float* d_dword, float* d_dpos, int32_t B, int32_t L, int32_t H, int32_t vocab, int32_t max_pos,
                       void* stream) {
  LDOT_REQUIRE(d_dx && d_ids && d_pos_ids && d_dword && d_dpos, "null pointer argument");
  return embed_scatter_run(d_dx, reinterpret_cast<const long long*>(d_ids), reinterpret_cast<const long long*>(d_pos_ids),
                           pos_batch_stride, d_dword, d_dpos, B, L, H, vocab, max_pos, stream);
}

int ldot_embed_image_pre(const float* d_lin, const float* d_box, const float* d_img_g, const float* d_img_b,
                         const float* d_pos_w, const float* d_pos_bias, const float* d_pos_g, const float* d_pos_b,
                         const float* d_type1, float* d_q, float* d_spre, int64_t rows, int32_t H, void* stream) {
  LDOT_REQUIRE(d_lin && d_box && d_img_g && d_img_b && d_pos_w && d_pos_bias && d_pos_g && d_pos_b && d_type1 && d_q &&
                   d_spre, "null pointer argument");
  EmbedImagePreParams p;
  p.lin = d_lin; p.box = d_box; p.img_g = d_img_g; p.img_b = d_img_b; p.pos_w = d_pos_w; p.pos_bias = d_pos_bias;
  p.pos_g = d_pos_g; p.pos_b = d_pos_b; p.type1 = d_type1; p.q = d_q; p.spre = d_spre; p.rows = rows;
  return embed_image_pre_run(p, H, stream);
}

int ldot_pos_wgrad(const float* d_dq, const float* d_box, int64_t rows, int32_t H, float* d_dw, void* stream) {
  LDOT_REQUIRE(d_dq && d_box && d_dw, "null pointer argument");
  return pos_wgrad_run(d_dq, d_box, rows, H, d_dw, stream);
}

int ldot_inbatch_nll_bwd(const float* d_scores, const int64_t* d_pos, int64_t bq, int64_t bc, const float* d_upstream,
                         int32_t reduction, void* d_dscores, int64_t ld_ds, int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_scores && d_pos && d_upstream && d_dscores, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  return nll_bwd_run(d_scores, reinterpret_cast<const long long*>(d_pos), bq, bc, d_upstream, reduction, d_dscores,
                     ld_ds, dtype, stream);
}

int ldot_sumsq(const float* d_g, int64_t n, float* d_out, void* stream) { return sumsq_run(d_g, n, d_out, stream); }

int ldot_adamw(float* d_p, const float* d_g, float* d_m, float* d_v, void* d_p16, int64_t n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int32_t step, const float* d_sumsq, float max_norm,
               int32_t dtype, void* stream) {
  LDOT_REQUIRE(step >= 1, "adamw: step counts from 1");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  AdamParams a;
  a.p = d_p; a.g = d_g; a.m = d_m; a.v = d_v; a.p16 = static_cast<uint16_t*>(d_p16); a.n = n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  a.bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), step)));
  a.sumsq = d_sumsq; a.max_norm = max_norm; a.fmt = dtype; a.hyper = nullptr;
  return adamw_run(a, stream);
}

// ---- fused training forms + CUDA-graph support (ABI 6) -----------------------------------------------------------
int ldot_linear_dropout(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias,
                        const void* d_residual, int64_t ldr, void* d_out, int64_t ldo, int64_t M, int32_t N, int32_t K,
                        int32_t dtype, float drop_p, uint64_t seed, int32_t site, void* stream) {
  LDOT_REQUIRE(d_a && d_w && d_out, "null pointer argument");
  LDOT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "linear_dropout: probability %f out of [0, 1)", drop_p);
  const DropKey drop = make_drop_key(drop_p, seed, site);
  return linear_run(d_a, lda, d_w, ldw, d_bias, d_residual, ldr, d_out, ldo, M, N, K, dtype, 0, 0, stream, &drop);
}

int ldot_linear_gelu_grad(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias, void* d_pre,
                         int64_t ld_pre, void* d_out, int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t dtype,
                         void* stream) {
  LDOT_REQUIRE(d_a && d_w && d_out && d_pre, "null pointer argument");
  return linear_run(d_a, lda, d_w, ldw, d_bias, nullptr, 0, d_out, ldo, M, N, K, dtype, 1, 0, stream, nullptr, d_pre, ld_pre);
}

int ldot_layernorm_bwd_dropout(const void* d_dy, int64_t ld_dy, const void* d_x, int64_t ld_x, const float* d_gamma,
                               void* d_dx, void* d_dx_masked, int64_t ld_dx, float* d_dgamma, float* d_dbeta,
                               float* d_dxsum, int64_t rows, int32_t H, float drop_p, uint64_t seed, int32_t site,
                               int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_dy && d_x && d_gamma && d_dx && d_dx_masked && d_dgamma && d_dbeta, "null pointer argument");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(drop_p > 0.f && drop_p < 1.f, "layernorm_bwd_dropout: probability %f out of (0, 1)", drop_p);
  LDOT_REQUIRE(H <= 1024, "layernorm_bwd_dropout: hidden size %d > 1024", H);
  LnBwdParams p;
  p.dy = d_dy; p.ld_dy = ld_dy; p.dy_f32 = 0; p.x = d_x; p.ld_x = ld_x; p.x_f32 = 0; p.gamma = d_gamma;
  p.dx = d_dx; p.ld_dx = ld_dx; p.dx_f32 = 0; p.dgamma = d_dgamma; p.dbeta = d_dbeta; p.dxsum = d_dxsum;
  p.rows = rows; p.fmt = dtype; p.dx_masked = d_dx_masked; p.drop = make_drop_key(drop_p, seed, site);
  return ln_bwd_run(p, H, stream);
}

int ldot_dropout_epoch(const uint32_t* d_epoch) {
  drop_epoch_slot() = d_epoch;
  return kOk;
}

int ldot_adamw_dev(float* d_p, const float* d_g, float* d_m, float* d_v, void* d_p16, int64_t n, const float* d_hyper,
                   float beta1, float beta2, float eps, float weight_decay, const float* d_sumsq, float max_norm,
                   int32_t dtype, void* stream) {
  LDOT_REQUIRE(d_hyper, "adamw_dev: null d_hyper");
  LDOT_REQUIRE(dtype == 0 || dtype == 1, "dtype must be 0 (fp16) or 1 (bf16)");
  AdamParams a;
  a.p = d_p; a.g = d_g; a.m = d_m; a.v = d_v; a.p16 = static_cast<uint16_t*>(d_p16); a.n = n;
  a.lr = 0.f; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay; a.bc1 = 1.f; a.bc2_sqrt = 1.f;
  a.sumsq = d_sumsq; a.max_norm = max_norm; a.fmt = dtype; a.hyper = d_hyper;
  return adamw_run(a, stream);
}

}  // extern "C"
