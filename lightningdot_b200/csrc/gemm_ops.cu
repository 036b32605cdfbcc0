// Encoder Linear layers on the tcgen05 main loop:  out = act(A . W^T + bias) (+ residual)
//
// Replaces every nn.Linear on the tower path (uniter_model/model/layer.py:64-66,76-78,107,133,148;
// uniter_model/model/model.py:252; dvl/models/bi_encoder.py:83-88,138-143).  A = activations [M, K] (16-bit,
// K-major), W = nn.Linear weight [N, K] (16-bit, K-major, used as stored - no transpose), fp32 accumulate in TMEM,
// bias / erf-GELU / residual add fused in the epilogue, 16-bit or fp32 output written by TMA stores.
// The kernel is linear_tc.cuh; this file is its host side.
#include "linear_tc.cuh"
#include "linear_ln.cuh"
#include "linear_ln2.cuh"
#include "qkv_attn.cuh"
#include "host_common.h"
#include "prof.h"
#include <cstdlib>

namespace ldot {

// 256 x 256 identity matrices (fp16 / bf16 1.0 = 0x3C00 / 0x3F80): the B operand that lets the tensor core add the
// residual tile.  Module-scope device memory, filled on first use (the library never allocates).
__device__ uint16_t g_eye[2][kLinBN * kLinBN];

__global__ void eye_init_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kLinBN * kLinBN) {
    const bool diag = (i / kLinBN) == (i % kLinBN);
    g_eye[0][i] = diag ? 0x3C00 : 0;
    g_eye[1][i] = diag ? 0x3F80 : 0;
  }
}

static int eye_pointer(int fmt, cudaStream_t st, const uint16_t** out) {
  static const uint16_t* base = nullptr;
  if (!base) {
    void* p = nullptr;
    LDOT_CUDA(cudaGetSymbolAddress(&p, g_eye));
    eye_init_kernel<<<(kLinBN * kLinBN + 255) / 256, 256, 0, st>>>();
    LDOT_CHECK_LAUNCH();
    // later launches may run on other streams: make the one-time fill visible to all of them
    LDOT_CUDA(cudaStreamSynchronize(st));
    base = static_cast<const uint16_t*>(p);
  }
  *out = base + static_cast<size_t>(fmt) * kLinBN * kLinBN;
  return kOk;
}

template <int ACT, int OUT_F32, int CTAS, int AMN = 0, int BMN = 0, int RED = 0, int TRAIN = 0>
static int launch_linear(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const LinSched& s,
                         const LinParams& p, int sms, cudaStream_t st) {
  auto kern = linear_tc_kernel<ACT, OUT_F32, CTAS, AMN, BMN, RED, TRAIN>;
  using SM = LinSmemT<CTAS, lin_epi_warps(ACT, TRAIN)>;
  static bool configured = false;  // (one device per process)
  static int max_groups = 0;       // resident CTAs (CTAS = 1) or CTA pairs (CTAS = 2)
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(lin_threads(ACT, TRAIN));
  cfg.dynamicSmemBytes = SM::kDynamic;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!configured) {
    LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kDynamic));
    if (CTAS == 2) {
      cfg.gridDim = dim3(CTAS);
      LDOT_CUDA(cudaOccupancyMaxActiveClusters(&max_groups, kern, &cfg));
      LDOT_REQUIRE(max_groups >= 1, "no resident CTA pair possible");
    } else {
      max_groups = sms;
    }
    configured = true;
  }
  const int groups = s.num_tiles < max_groups ? s.num_tiles : max_groups;
  cfg.gridDim = dim3(static_cast<unsigned>(groups * CTAS));
  LDOT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw, to, s, p));
  return kOk;
}

// CTA pairs (256 x 256 tiles, cta_group::2) pay off once the problem fills the machine several times over; small
// problems (projection head, [CLS]-only last layer, tests) keep 128 x 256 tiles so that more SMs get a tile.
// LDOT_LINEAR_CTAS=1|2 forces one form (measurement only).
static int linear_ctas(long long M, int N, int sms) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("LDOT_LINEAR_CTAS");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 1 || forced == 2) return forced;
  const long long tiles1 = ((M + kBM - 1) / kBM) * ((N + kLinBN - 1) / kLinBN);
  return tiles1 >= 4ll * sms ? 2 : 1;
}

int linear_run(const void* a, long long lda, const void* w, long long ldw, const float* bias, const void* residual,
               long long ldr, void* out, long long ldo, long long M, int N, int K, int fmt, int act, int out_f32,
               void* stream, const DropKey* drop = nullptr, void* pre = nullptr, long long ld_pre = 0) {
  LDOT_REQUIRE(M >= 1 && N >= 1 && K >= 8, "bad GEMM shape M=%lld N=%d K=%d", M, N, K);
  LDOT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "K, lda, ldw must be multiples of 8 elements");
  LDOT_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(act == 0 || act == 1, "act must be 0 (identity) or 1 (erf-GELU)");
  LDOT_REQUIRE(out_f32 ? (ldo % 4 == 0) : (ldo % 8 == 0), "ldo alignment");
  LDOT_REQUIRE(!residual || ldr % 8 == 0, "ldr alignment");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "out / residual / bias must be 16-byte aligned");
  LDOT_REQUIRE(M < (1ll << 31) - 256, "M too large");
  LDOT_REQUIRE(!drop || drop->thr == 0 || (act == 0 && N % 8 == 0), "the dropout epilogue needs act 0 and N % 8 == 0");
  LDOT_REQUIRE(!pre || (act == 1 && ld_pre % 8 == 0 && (reinterpret_cast<uintptr_t>(pre) & 15) == 0),
               "the pre-activation output needs act 1, ld_pre % 8 == 0 and a 16-byte aligned base");
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  const int ctas = linear_ctas(M, N, sms);
  CUtensorMap ta, tw, to;
  if (int e = make_tmap_kmajor_16b(&ta, a, M, K, static_cast<uint64_t>(lda) * 2, kBM)) return e;
  if (int e = make_tmap_kmajor_16b(&tw, w, N, K, static_cast<uint64_t>(ldw) * 2, kLinBN / ctas)) return e;
  const uint32_t elt = out_f32 ? 4 : 2;
  if (int e = make_tmap_store(&to, out, elt, M, N, static_cast<uint64_t>(ldo) * elt, 32, 32)) return e;
  LinSched s;
  s.m_tiles = static_cast<int>((M + kBM * ctas - 1) / (kBM * ctas));
  s.n_tiles = (N + kLinBN - 1) / kLinBN;
  s.num_tiles = s.m_tiles * s.n_tiles;
  s.k_blocks = (K + kBK - 1) / kBK;
  s.k_splits = 1;
  s.kb_per_split = s.k_blocks;
  s.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), kBM * ctas, kLinBN);
  LinParams p;
  p.bias = bias;
  p.residual = residual;
  p.ldr = ldr;
  p.M = M;
  p.N = N;
  p.fmt = fmt;
  p.drop = drop ? *drop : make_drop_key(0.f, 0, 0);
  p.pre = pre;
  p.ld_pre = ld_pre;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcLinear, st, 2.0 * M * static_cast<double>(N) * K,
                 (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 + static_cast<double>(M) * N * elt +
                     (residual ? static_cast<double>(M) * N * 2.0 : 0.0) + (pre ? static_cast<double>(M) * N * 2.0 : 0.0));
#define LDOT_LIN(ACT, F32) (ctas == 2 ? launch_linear<ACT, F32, 2>(ta, tw, to, s, p, sms, st) : launch_linear<ACT, F32, 1>(ta, tw, to, s, p, sms, st))
#define LDOT_LIN_TRAIN(ACT) (ctas == 2 ? launch_linear<ACT, 0, 2, 0, 0, 0, 1>(ta, tw, to, s, p, sms, st) : launch_linear<ACT, 0, 1, 0, 0, 0, 1>(ta, tw, to, s, p, sms, st))
  if (p.drop.thr != 0 || pre != nullptr) {
    LDOT_REQUIRE(!out_f32, "the training-forward epilogues write 16-bit output");
    return act == 1 ? LDOT_LIN_TRAIN(1) : LDOT_LIN_TRAIN(0);
  }
#undef LDOT_LIN_TRAIN
  if (act == 1) return out_f32 ? LDOT_LIN(1, 1) : LDOT_LIN(1, 0);
  return out_f32 ? LDOT_LIN(0, 1) : LDOT_LIN(0, 0);
#undef LDOT_LIN
}

// General form used by the training backward (include/ldot.h: ldot_gemm): either operand K-major or MN-major, optional
// accumulation into an fp32 output (split-K over the persistent grid), GELU-gradient / residual epilogues.
int gemm_run(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, const float* bias,
             const void* aux, long long ld_aux, void* out, long long ldo, long long M, int N, long long K, int fmt,
             int epi, int out_f32, int accumulate, void* stream) {
  LDOT_REQUIRE(M >= 1 && N >= 1 && K >= 1, "bad GEMM shape M=%lld N=%d K=%lld", M, N, K);
  LDOT_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(epi >= 0 && epi <= 4, "epi must be 0 (none), 1 (GELU), 2 (* GELU'(aux)), 3 (+ aux) or 4 (* aux)");
  LDOT_REQUIRE((epi >= 2) == (aux != nullptr), "aux is required by (and only by) epi 2 / 3 / 4");
  LDOT_REQUIRE(!accumulate || (out_f32 && epi == 0), "accumulate needs an fp32 output and no epilogue op");
  LDOT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "lda / ldb must be multiples of 8 elements");
  LDOT_REQUIRE(a_mn ? (M % 8 == 0) : (K % 8 == 0), "the contiguous extent of A must be a multiple of 8");
  LDOT_REQUIRE(b_mn ? (N % 8 == 0) : (K % 8 == 0), "the contiguous extent of B must be a multiple of 8");
  LDOT_REQUIRE(out_f32 ? (ldo % 4 == 0) : (ldo % 8 == 0), "ldo alignment");
  LDOT_REQUIRE(!aux || ld_aux % 8 == 0, "ld_aux alignment");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "out / aux / bias must be 16-byte aligned");
  LDOT_REQUIRE(M < (1ll << 31) - 256 && K < (1ll << 31) - 64, "M / K too large");
  const int variant = (a_mn ? 2 : 0) | (b_mn ? 1 : 0);
  if (variant == 0 && !accumulate && epi != 2 && epi != 4)
    return linear_run(a, lda, b, ldb, bias, epi == 3 ? aux : nullptr, ld_aux, out, ldo, M, N, static_cast<int>(K), fmt,
                      epi == 1 ? 1 : 0, out_f32, stream);
  // the forms the backward pass uses: dgrad (A K-major, B MN-major; 16-bit or fp32 out) and wgrad (both MN-major,
  // accumulating fp32)
  LDOT_REQUIRE((variant == 1 && !accumulate && epi != 1) || (variant == 3 && accumulate),
               "unsupported operand layout / epilogue combination (a_mn=%d b_mn=%d epi=%d accumulate=%d)", a_mn, b_mn,
               epi, accumulate);
  LDOT_REQUIRE(!((epi == 2 || epi == 4) && out_f32), "the GELU-gradient / product epilogues write 16-bit output");
  int sms = 0;
  if (int e = device_sm_count(&sms)) return e;
  // (wgrad tiles are few and K-split: CTA pairs would only halve the number of work items)
  const int ctas = variant == 3 ? 1 : linear_ctas(M, N, sms);
  CUtensorMap ta, tb, to;
  if (a_mn) {
    if (int e = make_tmap_kmajor_16b(&ta, a, K, M, static_cast<uint64_t>(lda) * 2, kBK)) return e;
  } else {
    if (int e = make_tmap_kmajor_16b(&ta, a, M, K, static_cast<uint64_t>(lda) * 2, kBM)) return e;
  }
  if (b_mn) {
    if (int e = make_tmap_kmajor_16b(&tb, b, K, N, static_cast<uint64_t>(ldb) * 2, kBK)) return e;
  } else {
    if (int e = make_tmap_kmajor_16b(&tb, b, N, K, static_cast<uint64_t>(ldb) * 2, kLinBN / ctas)) return e;
  }
  const uint32_t elt = out_f32 ? 4 : 2;
  if (int e = make_tmap_store(&to, out, elt, M, N, static_cast<uint64_t>(ldo) * elt, 32, 32, accumulate != 0)) return e;
  LinSched s;
  s.m_tiles = static_cast<int>((M + kBM * ctas - 1) / (kBM * ctas));
  s.n_tiles = (N + kLinBN - 1) / kLinBN;
  s.k_blocks = static_cast<int>((K + kBK - 1) / kBK);
  s.k_splits = 1;
  s.kb_per_split = s.k_blocks;
  if (accumulate) {
    // enough K slices that every SM (pair) gets about two work items
    const int mn = s.m_tiles * s.n_tiles, groups = sms / ctas;
    int want = (2 * groups + mn - 1) / mn;
    want = want < 1 ? 1 : (want > s.k_blocks ? s.k_blocks : want);
    s.kb_per_split = (s.k_blocks + want - 1) / want;
    s.k_splits = (s.k_blocks + s.kb_per_split - 1) / s.kb_per_split;
  }
  s.num_tiles = s.m_tiles * s.n_tiles * s.k_splits;
  s.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), kBM * ctas, kLinBN, a_mn ? 1u : 0u, b_mn ? 1u : 0u);
  LinParams p;
  p.bias = bias;
  p.residual = aux;
  p.ldr = ld_aux;
  p.M = M;
  p.N = N;
  p.fmt = fmt;
  p.drop = make_drop_key(0.f, 0, 0);
  p.pre = nullptr;
  p.ld_pre = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KernelScope ks(kKcLinear, st, 2.0 * M * static_cast<double>(N) * K,
                 (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 + static_cast<double>(M) * N * elt +
                     (aux ? static_cast<double>(M) * N * 2.0 : 0.0));
#define LDOT_G(ACT, F32, AMN, BMN, RED)                                                     \
  (ctas == 2 ? launch_linear<ACT, F32, 2, AMN, BMN, RED>(ta, tb, to, s, p, sms, st)        \
             : launch_linear<ACT, F32, 1, AMN, BMN, RED>(ta, tb, to, s, p, sms, st))
  if (variant == 3) return launch_linear<0, 1, 1, 1, 1, 1>(ta, tb, to, s, p, sms, st);
  if (epi == 2) return LDOT_G(2, 0, 0, 1, 0);
  if (epi == 4) return LDOT_G(3, 0, 0, 1, 0);
  return out_f32 ? LDOT_G(0, 1, 0, 1, 0) : LDOT_G(0, 0, 0, 1, 0);
#undef LDOT_G
}

// out = LayerNorm(A . W^T + bias + residual) * gamma + beta, 16-bit output; a cluster of ceil(N / 256) CTAs per row block
// CTA-pair form (linear_ln2.cuh): N == 768 with a residual and enough rows to fill the machine.  LDOT_LN_PAIR=0 disables it.
static int linear_ln_pair_run(const void* a, long long lda, const void* w, long long ldw, const float* bias,
                              const void* residual, long long ldr, const float* gamma, const float* beta, void* out,
                              long long ldo, long long M, int N, int K, int fmt, cudaStream_t st) {
  CUtensorMap ta, tw, to, tr, te;
  const uint16_t* eye = nullptr;
  if (int e = eye_pointer(fmt, st, &eye)) return e;
  if (int e = make_tmap_kmajor_16b(&ta, a, M, K, static_cast<uint64_t>(lda) * 2, kBM)) return e;
  if (int e = make_tmap_kmajor_16b(&tw, w, N, K, static_cast<uint64_t>(ldw) * 2, kLinBN / 2)) return e;
  if (int e = make_tmap_store(&to, out, 2, M, N, static_cast<uint64_t>(ldo) * 2, 32, 32)) return e;
  if (int e = make_tmap_kmajor_16b(&tr, residual, M, N, static_cast<uint64_t>(ldr) * 2, kBM)) return e;
  if (int e = make_tmap_kmajor_16b(&te, eye, kLinBN, kLinBN, kLinBN * 2, kBK / 2)) return e;   // 32-row halves of I_64
  static bool configured = false;
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 6;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kLinThreads);
  cfg.dynamicSmemBytes = Ln2Smem::kDynamic;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!configured) {
    LDOT_CUDA(cudaFuncSetAttribute(linear_ln2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Ln2Smem::kDynamic));
    cfg.gridDim = dim3(6);
    LDOT_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, linear_ln2_kernel, &cfg));
    LDOT_REQUIRE(max_clusters >= 1, "no resident cluster of 6 CTAs possible");
    configured = true;
  }
  LnSched s;
  s.m_tiles = static_cast<int>((M + 2 * kBM - 1) / (2 * kBM));   // 256-row blocks, one per cluster pass
  s.k_blocks = (K + kBK - 1) / kBK;
  s.cluster = 3;
  s.res_blocks = kLinBN / kBK;
  s.num_clusters = s.m_tiles < max_clusters ? s.m_tiles : max_clusters;
  s.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), 2 * kBM, kLinBN);
  s.idesc_res = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), 2 * kBM, kBK);
  LnParams p;
  p.bias = bias;
  p.gamma = gamma;
  p.beta = beta;
  p.M = M;
  p.N = N;
  p.fmt = fmt;
  cfg.gridDim = dim3(static_cast<unsigned>(s.num_clusters * 6));
  KernelScope ks(kKcLinear, st, 2.0 * M * static_cast<double>(N) * K,
                 (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 + static_cast<double>(M) * N * 4.0);
  LDOT_CUDA(cudaLaunchKernelEx(&cfg, linear_ln2_kernel, ta, tw, tr, te, to, s, p));
  return kOk;
}

int linear_ln_run(const void* a, long long lda, const void* w, long long ldw, const float* bias, const void* residual,
                  long long ldr, const float* gamma, const float* beta, void* out, long long ldo, long long M, int N,
                  int K, int fmt, void* stream) {
  LDOT_REQUIRE(M >= 1 && N >= 32 && K >= 8, "bad GEMM shape M=%lld N=%d K=%d", M, N, K);
  LDOT_REQUIRE(N % 32 == 0 && N <= kLnMaxCluster * kLinBN, "linear+LayerNorm needs N %% 32 == 0 and N <= %d, got %d",
               kLnMaxCluster * kLinBN, N);
  LDOT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0, "K, lda, ldw, ldo must be multiples of 8");
  LDOT_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(gamma && beta, "gamma / beta are required");
  LDOT_REQUIRE(!residual || ldr % 8 == 0, "ldr alignment");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(bias) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(beta) & 15) == 0,
               "out / residual / bias / gamma / beta must be 16-byte aligned");
  LDOT_REQUIRE(M < (1ll << 31) - 128, "M too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = (N + kLinBN - 1) / kLinBN;
  {
    static int pair = -1;
    if (pair < 0) {
      const char* e = getenv("LDOT_LN_PAIR");
      pair = e ? atoi(e) : 1;
    }
    if (pair && N == 3 * kLinBN && residual != nullptr && M >= 64 * 2 * kBM)
      return linear_ln_pair_run(a, lda, w, ldw, bias, residual, ldr, gamma, beta, out, ldo, M, N, K, fmt, st);
  }
  CUtensorMap ta, tw, to;
  if (int e = make_tmap_kmajor_16b(&ta, a, M, K, static_cast<uint64_t>(lda) * 2, kBM)) return e;
  if (int e = make_tmap_kmajor_16b(&tw, w, N, K, static_cast<uint64_t>(ldw) * 2, kLinBN)) return e;
  if (int e = make_tmap_store(&to, out, 2, M, N, static_cast<uint64_t>(ldo) * 2, 32, 32)) return e;
  CUtensorMap tr = ta, te = tw;  // (unused unless there is a residual)
  if (residual) {
    const uint16_t* eye = nullptr;
    if (int e = eye_pointer(fmt, st, &eye)) return e;
    if (int e = make_tmap_kmajor_16b(&tr, residual, M, N, static_cast<uint64_t>(ldr) * 2, kBM)) return e;
    if (int e = make_tmap_kmajor_16b(&te, eye, kLinBN, kLinBN, kLinBN * 2, kBK)) return e;   // 64 x 64 box: I_64
  }

  static bool configured = false;
  static int max_clusters[kLnMaxCluster + 1] = {0};
  if (!configured) {
    LDOT_CUDA(cudaFuncSetAttribute(linear_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LnSmem::kDynamic));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(C);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kLinThreads);
  cfg.dynamicSmemBytes = LnSmem::kDynamic;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters[C] == 0) {
    // how many clusters of C CTAs (one CTA per SM) the device keeps resident: the persistent grid must not exceed it
    cfg.gridDim = dim3(static_cast<unsigned>(C));
    int n = 0;
    LDOT_CUDA(cudaOccupancyMaxActiveClusters(&n, linear_ln_kernel, &cfg));
    LDOT_REQUIRE(n >= 1, "no resident cluster of %d CTAs possible", C);
    max_clusters[C] = n;
  }
  LnSched s;
  s.m_tiles = static_cast<int>((M + kBM - 1) / kBM);
  s.k_blocks = (K + kBK - 1) / kBK;
  s.cluster = C;
  s.res_blocks = residual ? kLinBN / kBK : 0;
  s.num_clusters = s.m_tiles < max_clusters[C] ? s.m_tiles : max_clusters[C];
  s.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), kBM, kLinBN);
  s.idesc_res = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), kBM, kBK);
  LnParams p;
  p.bias = bias;
  p.gamma = gamma;
  p.beta = beta;
  p.M = M;
  p.N = N;
  p.fmt = fmt;
  cfg.gridDim = dim3(static_cast<unsigned>(s.num_clusters * C));
  KernelScope ks(kKcLinear, st, 2.0 * M * static_cast<double>(N) * K,
                 (static_cast<double>(M) * K + static_cast<double>(N) * K) * 2.0 + static_cast<double>(M) * N * 2.0 +
                     (residual ? static_cast<double>(M) * N * 2.0 : 0.0));
  LDOT_CUDA(cudaLaunchKernelEx(&cfg, linear_ln_kernel, ta, tw, tr, te, to, s, p));
  return kOk;
}


// ---------------------------------------------------------------------------------------- fused QKV projection + attention
template <int SPAD, int FMT>
static int launch_qkv_attn(const CUtensorMap& ta, const CUtensorMap& tw, const QaSched& s, const QaParams& p, cudaStream_t st) {
  auto kern = qkv_attn_kernel<SPAD, FMT>;
  static bool configured = false;
  static int max_pairs = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kQaThreads);
  cfg.dynamicSmemBytes = QaSmem::kDynamic;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!configured) {
    LDOT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QaSmem::kDynamic));
    cfg.gridDim = dim3(2);
    LDOT_CUDA(cudaOccupancyMaxActiveClusters(&max_pairs, kern, &cfg));
    LDOT_REQUIRE(max_pairs >= 1, "no resident CTA pair possible");
    configured = true;
  }
  const int pairs = s.num_tiles < max_pairs ? s.num_tiles : max_pairs;
  cfg.gridDim = dim3(static_cast<unsigned>(pairs * 2));
  LDOT_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tw, s, p));
  return kOk;
}

// x [B * S, K] (16-bit) -> ctx [B * S, H]: BertSelfAttention (Q | K | V projection with the stacked [3 H, K] weight, scaled
// dot-product attention with the additive mask, per head) without the [B * S, 3 H] intermediate in HBM.
int qkv_attention_run(const void* x, long long ldx, const void* w, long long ldw, const float* bias, const long long* mask,
                      void* ctx, int B, int S, int H, int heads, int K, int fmt, void* stream) {
  LDOT_REQUIRE(B >= 1 && S >= 1 && S <= 128, "qkv_attention: bad shape B=%d S=%d (S <= 128)", B, S);
  LDOT_REQUIRE(H == heads * kHeadDim, "qkv_attention: hidden %d must be heads (%d) x 64", H, heads);
  LDOT_REQUIRE(K >= 8 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "K, ldx, ldw must be multiples of 8 elements");
  LDOT_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (fp16) or 1 (bf16)");
  LDOT_REQUIRE(bias != nullptr, "qkv_attention: the projection bias is required");
  LDOT_REQUIRE((reinterpret_cast<uintptr_t>(ctx) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
               "ctx / bias must be 16-byte aligned");
  const long long T = static_cast<long long>(B) * S;
  LDOT_REQUIRE(T < (1ll << 31) - 256, "too many tokens");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap ta, tw;
  if (int e = make_tmap_kmajor_16b(&ta, x, static_cast<uint64_t>(T), K, static_cast<uint64_t>(ldx) * 2, kBM)) return e;
  if (int e = make_tmap_kmajor_16b(&tw, w, static_cast<uint64_t>(3) * H, K, static_cast<uint64_t>(ldw) * 2, 32)) return e;
  QaSched s;
  s.seq_per_tile = kBM / S;
  s.rows_per_tile = s.seq_per_tile * S;
  s.m_tiles = (B + s.seq_per_tile - 1) / s.seq_per_tile;
  s.heads = heads;
  s.num_tiles = ((s.m_tiles + 1) / 2) * heads;
  s.k_blocks = (K + kBK - 1) / kBK;
  s.idesc = ptx::make_idesc_f16(static_cast<uint32_t>(fmt), kBM * 2, kQaBN);
  static int yield_mma = -1;
  if (yield_mma < 0) {
    const char* e = getenv("LDOT_QA_YIELD");   // (measurement switch)
    yield_mma = e ? atoi(e) : 2;
  }
  s.yield_mma = yield_mma;
  QaParams p;
  p.bias = bias;
  p.mask = mask;
  p.ctx = static_cast<uint16_t*>(ctx);
  p.B = B;
  p.S = S;
  p.H = H;
  // algorithmic work: the projection GEMM + the two attention contractions; bytes: x + W + ctx (no Q | K | V round trip)
  KernelScope ks(kKcQkvAttn, st, 2.0 * T * 3.0 * H * K + 4.0 * T * S * H,
                 (static_cast<double>(T) * K + 3.0 * H * K + static_cast<double>(T) * H) * 2.0);
  const int spad = S <= 32 ? 32 : S <= 48 ? 48 : S <= 64 ? 64 : S <= 96 ? 96 : 128;
#define LDOT_QA(SP) (fmt == 1 ? launch_qkv_attn<SP, 1>(ta, tw, s, p, st) : launch_qkv_attn<SP, 0>(ta, tw, s, p, st))
  switch (spad) {
    case 32: return LDOT_QA(32);
    case 48: return LDOT_QA(48);
    case 64: return LDOT_QA(64);
    case 96: return LDOT_QA(96);
    default: return LDOT_QA(128);
  }
#undef LDOT_QA
}

}  // namespace ldot
