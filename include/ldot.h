/* libldot_sm100a - C ABI of the B200-native LightningDOT retrieval hot path.
 *
 * Conventions (all entry points):
 *   - return 0 on success, < 0 on error; ldot_last_error() returns a thread-local description of the last failure;
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device; `stream` is a cudaStream_t (NULL =
 *     legacy default stream); calls only enqueue work - they never synchronise and never allocate: scratch memory
 *     is passed in by the caller and sized with the matching *_workspace_bytes() query;
 *   - row-major, contiguous matrices; row ids are int64 like faiss labels.
 *
 * The reference has no FFI for this path (it is pure Python over third-party wheels); each entry point cites the
 * reference call it replaces.  See INTEGRATION.md for the ctypes binding the reference's Python would add.
 */
#ifndef LDOT_H_
#define LDOT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDOT_ABI_VERSION 6

#define LDOT_OK 0
#define LDOT_ERR_ARG (-1)
#define LDOT_ERR_CUDA (-2)
#define LDOT_ERR_WORKSPACE (-3)
#define LDOT_ERR_ARCH (-4)

/* 16-bit storage type of the coarse (tensor-core) copy of queries and index */
#define LDOT_COARSE_FP16 0
#define LDOT_COARSE_BF16 1

int ldot_abi_version(void);
const char* ldot_last_error(void);
/* 0 when the current device is an sm_100 part the kernels can run on, LDOT_ERR_ARCH / LDOT_ERR_CUDA otherwise */
int ldot_device_check(void);

/* ---- launch accounting (measurement support; no reference counterpart) ------------------------------------------
 * The library counts every kernel it launches per kernel class (always) and, between ldot_prof_enable(1) and
 * ldot_prof_enable(0), brackets each launch with CUDA events on the launching stream.  ldot_prof_read waits for
 * the recorded events and returns, per class: launches, launches that were timed, summed device milliseconds, and
 * the algorithmic FLOPs / bytes of those launches (DESIGN.md states the per-unit figures).  These three calls are
 * the only entry points that synchronise (on their own events only).                                             */
int ldot_prof_num_classes(void);
const char* ldot_prof_class_name(int32_t cls);
int ldot_prof_enable(int32_t on);
int ldot_prof_reset(void);
int ldot_prof_read(int32_t n_classes, int64_t* launches, int64_t* timed_launches, double* ms, double* flops,
                   double* bytes);

/* ---- index build: replaces faiss.IndexFlatIP.add (dvl/indexer/faiss_indexers.py:77) -------------------------
 * d_x      [n, d] fp32 master copy of the index (kept by the caller; the search reads it for exact rescoring)
 * d_x16    [n, d] out: centred 16-bit copy (x - mean row) for the tensor-core pass
 * d_mu     [d]    out: the centring vector (all zeros when center == 0)
 * d_xstats [2]    out: max_j |(x_j - mu) - x16_j|_2 and max_j |x16_j|_2 (inputs of the search certificate)     */
size_t ldot_index_prepare_workspace_bytes(int64_t n, int32_t d);
int ldot_index_prepare(const float* d_x, int64_t n, int32_t d, int32_t coarse_dtype, int32_t center, void* d_x16,
                       float* d_mu, float* d_xstats, void* d_ws, size_t ws_bytes, void* stream);

/* ---- search: replaces faiss.IndexFlatIP.search (dvl/indexer/faiss_indexers.py:83) ----------------------------
 * Exact inner-product top-k of every query against every index row, ranked (score desc, row id asc).
 * d_q          [nq, d] fp32 queries
 * coarse_k     candidates kept by the tensor-core pass per query (0 = automatic, >= k)
 * id_offset    added to every returned row id (row offset of this shard in a row-sharded index)
 * d_out_scores [nq, k] fp32, d_out_idx [nq, k] int64 (-1 / -FLT_MAX past the end of a short index)
 * d_out_flags  [nq] int32: 0 = result proven exact; 1 = certificate failed - the caller must re-run that query
 *              through ldot_flatip_exact (DenseFlatIndexer.search_knn does)
 * d_out_flag_count [1] int32 number of flagged queries: device or PINNED HOST memory (UVA copy), may be NULL     */
size_t ldot_flatip_search_workspace_bytes(int64_t nq, int64_t n, int32_t d, int32_t k, int32_t coarse_k);
int ldot_flatip_search(const float* d_q, int64_t nq, const float* d_x, const void* d_x16, const float* d_mu,
                       const float* d_xstats, int64_t n, int32_t d, int32_t k, int32_t coarse_k,
                       int32_t coarse_dtype, int64_t id_offset, float* d_out_scores, int64_t* d_out_idx,
                       int32_t* d_out_flags, int32_t* d_out_flag_count, void* d_ws, size_t ws_bytes, void* stream);

/* Exhaustive fp64-accumulated scan (no tensor cores): the fallback for flagged queries; same outputs/ranking. */
/* Row-SHARDED search in two phases over the same workspace (lightningdot_b200/sharded.py): every shard runs phase 1,
 * the shards take the MINIMUM of their d_bound vectors (one all-reduce of nq floats), every shard runs phase 2.
 *   phase 1  query prepare .. candidate selection; d_bound[q] = a score that at least bound_m rows of THIS shard reach
 *            (bound_m-th best coarse score + q.mu - error bound; -inf if the shard holds fewer rows).  With
 *            world * bound_m >= k the minimum over the shards is a lower bound tau[q] of the global k-th best score.
 *   phase 2  exact rescoring of the candidates that can reach d_tau[q], ranking, certificate (a query is certified when
 *            the usual local condition holds OR no row outside the candidate list can reach tau); fewer than k rows may
 *            survive: the tail of the list is (-FLT_MAX, -1), which ldot_topk_merge skips.
 *   phase 0  the whole search in one call (== ldot_flatip_search; d_bound / d_tau unused).
 * Arguments as ldot_flatip_search; the workspace must not be touched between the phases.                            */
int ldot_flatip_search_phase(const float* d_q, int64_t nq, const float* d_x, const void* d_x16, const float* d_mu,
                             const float* d_xstats, int64_t n, int32_t d, int32_t k, int32_t coarse_k,
                             int32_t coarse_dtype, int64_t id_offset, float* d_out_scores, int64_t* d_out_idx,
                             int32_t* d_out_flags, int32_t* d_out_flag_count, void* d_ws, size_t ws_bytes, int32_t phase,
                             int32_t bound_m, float* d_bound, const float* d_tau, void* stream);

size_t ldot_flatip_exact_workspace_bytes(int64_t n);
int ldot_flatip_exact(const float* d_q, int64_t nq, const float* d_x, int64_t n, int32_t d, int32_t k,
                      int64_t id_offset, float* d_out_scores, int64_t* d_out_idx, void* d_ws, size_t ws_bytes,
                      void* stream);

/* Merge the exact top-k lists of `world` index shards (after the NCCL exchange):
 * d_scores [world, nq, k], d_idx [world, nq, k] (global ids; every list ranked as ldot_flatip_search returns it:
 * (score desc, id asc), label -1 entries last) -> d_out_* [nq, k], same ranking rule.  shard_stride_* = elements between
 * the lists of consecutive shards (0 = dense, nq * k): lets scores and ids of a shard travel in one packed buffer.     */
int ldot_topk_merge(const float* d_scores, const int64_t* d_idx, int32_t world, int64_t nq, int32_t k,
                    int64_t shard_stride_scores, int64_t shard_stride_idx, float* d_out_scores, int64_t* d_out_idx,
                    void* stream);

/* ---- encoder Linear: replaces nn.Linear (+ fused GELU / residual) on the tower path ---------------------------
 * (uniter_model/model/layer.py:76-78,107,133,148; model.py:252; dvl/models/bi_encoder.py:83-88,138-143)
 * out[M, N] = act(A[M, K] . W[N, K]^T + bias[N]) (+ residual[M, N]);  A / W / residual: 16-bit of `dtype`
 * (LDOT_COARSE_FP16 / LDOT_COARSE_BF16), fp32 accumulation; lda / ldw / ldr / ldo are row pitches in ELEMENTS;
 * act: 0 identity, 1 erf-GELU; out_f32: 1 = fp32 output, 0 = 16-bit output.  bias / residual may be NULL.        */
int ldot_linear(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias,
                const void* d_residual, int64_t ldr, void* d_out, int64_t ldo, int64_t M, int32_t N, int32_t K,
                int32_t dtype, int32_t act, int32_t out_f32, void* stream);

/* ---- Linear + residual + LayerNorm fused: replaces BertSelfOutput / BertOutput (uniter_model/model/layer.py:104-115,
 * 145-156):  out[M, N] = LayerNorm(A . W^T + bias + residual) * gamma + beta, eps 1e-12, 16-bit output of `dtype`.
 * N % 32 == 0 and N <= 768 (the 128-row block is shared by a cluster of ceil(N / 256) CTAs that exchange fp32 row
 * statistics through distributed shared memory); bias / residual may be NULL; gamma / beta fp32 [N].              */
int ldot_linear_ln(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias,
                   const void* d_residual, int64_t ldr, const float* d_gamma, const float* d_beta, void* d_out,
                   int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t dtype, void* stream);

/* ---- the rest of the tower arithmetic (16-bit activations of `dtype`, fp32 statistics) -------------------------
 * ldot_layernorm   out[rows, H] = LayerNorm(in[rows, H]) * gamma + beta, eps 1e-12 (layer.py:108-115,149-156);
 *                  `in` is fp32 when in_f32 != 0, else 16-bit; ld_* are row pitches in elements; H = 256 * {1,2,3,4,6,8}
 * ldot_embed_text  out[b * out_seq + l, :] = LN(word[ids[b, l]] + pos[pos_ids[b * pos_batch_stride + l]] + type0)
 *                  (model.py:233-246); ids / pos_ids int64; tables 16-bit [vocab, H] / [max_pos, H] / [H]
 * ldot_embed_image out[b * out_seq + row_offset + r, :] = LN(LN_img(lin[b, r]) + LN_pos(W_pos box[b, r] + b_pos) + type1)
 *                  (model.py:262-273,328-336); lin = img_linear(feat) + bias as fp32 [B * R, H] (from ldot_linear)
 * ldot_attention   ctx = softmax(Q K^T / 8 + (1 - mask) * -10000) V per (sequence, head), head dim 64
 *                  (layer.py:80-101, model.py:362-365); qkv [B * S, 3 H] = Q | K | V; mask int64 [B, S]; S <= 128;
 *                  only the first q_rows query positions of each sequence are computed: ctx is [B * q_rows, H]
 *                  (q_rows = S: the whole layer; q_rows = 1: the last layer when only the [CLS] row is read,
 *                  dvl/models/bi_encoder.py:120,188)
 * ldot_cast_f32    fp32 -> 16-bit, n elements (n % 8 == 0)                                                          */
int ldot_layernorm(const void* d_in, int64_t ld_in, int32_t in_f32, const float* d_gamma, const float* d_beta,
                   void* d_out, int64_t ld_out, int64_t rows, int32_t H, int32_t dtype, void* stream);
int ldot_embed_text(const int64_t* d_ids, const int64_t* d_pos_ids, int64_t pos_batch_stride, const void* d_word,
                    const void* d_pos, const void* d_type0, const float* d_gamma, const float* d_beta, void* d_out,
                    int32_t B, int32_t L, int32_t out_seq, int32_t H, int32_t vocab, int32_t max_pos, int32_t dtype,
                    void* stream);
int ldot_embed_image(const float* d_lin, const float* d_box, const float* d_img_g, const float* d_img_b,
                     const float* d_pos_w, const float* d_pos_bias, const float* d_pos_g, const float* d_pos_b,
                     const float* d_type1, const float* d_ln_g, const float* d_ln_b, void* d_out, int32_t B, int32_t R,
                     int32_t out_seq, int32_t row_offset, int32_t H, int32_t dtype, void* stream);
int ldot_attention(const void* d_qkv, const int64_t* d_mask, void* d_ctx, int32_t B, int32_t S, int32_t H,
                   int32_t heads, int32_t q_rows, int32_t dtype, void* stream);
int ldot_cast_f32(const float* d_in, void* d_out, int64_t n, int32_t dtype, void* stream);
/* ldot_qkv_attention  BertSelfAttention in one kernel (uniter_model/model/layer.py:60-101): d_ctx [B * S, H] =
 *                     attention over Q | K | V = d_x [B * S, K] . d_w [3 H, K]^T + d_bias [3 H] (the query / key / value
 *                     nn.Linear weights stacked in that order), additive mask from d_mask int64 [B, S] (1 = attend),
 *                     heads of 64.  Replaces ldot_linear (N = 3 H) + ldot_attention (q_rows = S): the [B * S, 3 H]
 *                     projection never reaches HBM.  S <= 128.                                                          */
int ldot_qkv_attention(const void* d_x, int64_t ldx, const void* d_w, int64_t ldw, const float* d_bias,
                       const int64_t* d_mask, void* d_ctx, int32_t B, int32_t S, int32_t H, int32_t heads, int32_t K,
                       int32_t dtype, void* stream);
/* Training-mode dropout of the towers (bi_encoder.py:97-99 -> hidden_dropout_prob, attention_probs_dropout_prob): masks
 * are a counter-based function of (seed, site, element index) - csrc/dropout.cuh - regenerated by the backward kernels.
 * ldot_attention_train  ldot_attention (q_rows = S) with dropout of the attention probabilities (layer.py:93)
 * ldot_dropout          d_out = dropout(d_x) (+ d_res): 16-bit [rows, cols], row pitch ld, in place allowed (layer.py:113,154) */
int ldot_attention_train(const void* d_qkv, const int64_t* d_mask, void* d_ctx, int32_t B, int32_t S, int32_t H,
                         int32_t heads, float drop_p, uint64_t seed, int32_t site, int32_t dtype, void* stream);
int ldot_dropout(const void* d_x, const void* d_res, void* d_out, int64_t rows, int32_t cols, int64_t ld, float p,
                 uint64_t seed, int32_t site, int32_t dtype, void* stream);

/* ---- in-batch-negative NLL: replaces BiEncoderNllLoss.calc (dvl/models/bi_encoder.py:615-656) -------------------
 * ldot_split16     fp32 [rows, K] -> fp16 [rows, 3 K] hi/lo split, side 0: [hi | lo | hi], side 1: [hi | hi | lo];
 *                  ldot_linear(split(Q, 0), split(C, 1), K = 3 K, fp32 out) then yields Q . C^T to ~2^-22 relative
 *                  (dot_product_scores, bi_encoder.py:54-68) on the 16-bit tensor cores
 * ldot_inbatch_nll scores = d_scores (mixed with d_scores_cap: (1 - w) s + w s_cap when d_scores_cap != NULL,
 *                  bi_encoder.py:625-627) -> d_scores_out [bq, bc] (may alias d_scores);
 *                  d_row_loss[i] = logsumexp_j s[i, j] - s[i, pos[i]]; d_loss = mean (reduction 0) or sum (1);
 *                  d_correct = #{ i : argmax_j s[i, j] == pos[i] } (first maximal index); d_row_correct [bq] scratch */
int ldot_split16(const float* d_in, int64_t rows, int32_t K, int32_t side, void* d_out, void* stream);
int ldot_inbatch_nll(const float* d_scores, const float* d_scores_cap, float cap_weight, const int64_t* d_pos,
                     int64_t bq, int64_t bc, int32_t reduction, float* d_scores_out, float* d_row_loss,
                     int32_t* d_row_correct, float* d_loss, int64_t* d_correct, void* stream);

/* ---- training step: the backward of train_itm.py:191-289 through the towers and the loss ---------
 * The reference gets all of this from torch autograd + apex; these entry points are what the autograd Functions of
 * lightningdot_b200/training.py call.  Activations and activation gradients are 16-bit of `dtype`, parameter gradients
 * fp32 and ACCUMULATED (+=) like torch's .grad.
 *
 * ldot_gemm          out[M, N] (+)= op(A) . op(B)^T on the tcgen05 kernel with either operand K-major (a_mn / b_mn = 0:
 *                    row-major [M, K] / [N, K]) or MN-major (1: row-major [K, M] / [K, N], used as stored):
 *                      forward  x W^T          a_mn 0, b_mn 0   (== ldot_linear)
 *                      dgrad    dX = dY W      a_mn 0, b_mn 1   A = dY [M, K = out], B = W [K = out, N = in]
 *                      wgrad    dW += dY^T X   a_mn 1, b_mn 1   A = dY [K = tokens, M = out], B = X [K = tokens, N = in],
 *                                                               accumulate = 1, fp32 out, split-K over the grid with
 *                                                               TMA reduce-add stores
 *                    epi: 0 none, 1 erf-GELU, 2 out = acc * GELU'(aux[m, n]), 3 out = acc + aux[m, n], 4 out = acc * aux[m, n]
 *                    (aux 16-bit)
 * ldot_layernorm_bwd dx = LayerNorm backward of y = LN(x) * gamma + beta given dy; dgamma / dbeta [H] += ; when
 *                    d_dxsum != NULL also dxsum[H] += sum_rows dx (the bias gradient of the Linear that produced x).
 *                    *_f32 flags: the tensor is fp32 instead of 16-bit.  H = 256 * {1,2,3,4,6}
 * ldot_attention_bwd d_dqkv [B * S, 3 H] = backward of ldot_attention (q_rows = S) given d_dctx [B * S, H]; the
 *                    probabilities are recomputed from the saved d_qkv, d_ctx is the saved forward output; drop_p / seed /
 *                    site = the attention dropout of the matching ldot_attention_train call (0 = none)
 * ldot_gelu / ldot_gelu_bwd      elementwise erf-GELU (16-bit) and dx = dy * GELU'(x); n % 8 == 0
 * ldot_colsum16      d_out[N] += column sums of a 16-bit [rows, N] matrix (bias gradients)
 * ldot_embed_text_sum / ldot_embed_scatter   text-embedding backward (uniter_model/model/model.py:233-246): the fp32
 *                    LayerNorm input word + pos + type0 recomputed; dword[ids] += dx (row 0 = padding_idx skipped),
 *                    dpos[pos_ids] += dx
 * ldot_embed_image_pre / ldot_pos_wgrad      image-embedding backward (model.py:262-273): the fp32 LayerNorm inputs
 *                    q = pos_linear(box) and spre = LN_img(lin) + LN_pos(q) + type1 recomputed; dW_pos[H, 7] += dq^T box
 * ldot_inbatch_nll_bwd   d_dscores [bq, ld_ds] 16-bit = upstream * (softmax(scores) - onehot(pos)) (/ bq for the mean
 *                    reduction), columns bc .. ld_ds - 1 zero (bi_encoder.py:632-640 under autograd)
 * ldot_sumsq         d_out[0] += sum g^2 (gradient norm, train_itm.py:262-267)
 * ldot_adamw         torch.optim.AdamW step over one flat fp32 tensor (bi_encoder.py:566-576), `step` counted from 1;
 *                    d_sumsq != NULL: gradients are first scaled by min(1, max_norm / (sqrt(*d_sumsq) + 1e-6));
 *                    d_p16 != NULL: the updated parameter is also written as 16-bit of `dtype`                      */
int ldot_gemm(const void* d_a, int64_t lda, int32_t a_mn, const void* d_b, int64_t ldb, int32_t b_mn,
              const float* d_bias, const void* d_aux, int64_t ld_aux, void* d_out, int64_t ldo, int64_t M, int32_t N,
              int64_t K, int32_t dtype, int32_t epi, int32_t out_f32, int32_t accumulate, void* stream);
int ldot_layernorm_bwd(const void* d_dy, int64_t ld_dy, int32_t dy_f32, const void* d_x, int64_t ld_x, int32_t x_f32,
                       const float* d_gamma, void* d_dx, int64_t ld_dx, int32_t dx_f32, float* d_dgamma,
                       float* d_dbeta, float* d_dxsum, int64_t rows, int32_t H, int32_t dtype, void* stream);
int ldot_attention_bwd(const void* d_qkv, const int64_t* d_mask, const void* d_ctx, const void* d_dctx, void* d_dqkv,
                       int32_t B, int32_t S, int32_t H, int32_t heads, float drop_p, uint64_t seed, int32_t site,
                       int32_t dtype, void* stream);
int ldot_gelu(const void* d_x, void* d_out, int64_t n, int32_t dtype, void* stream);
int ldot_gelu_bwd(const void* d_x, const void* d_dy, void* d_dx, int64_t n, int32_t dtype, void* stream);
int ldot_colsum16(const void* d_in, int64_t ld, int64_t rows, int32_t N, float* d_out, int32_t dtype, void* stream);
int ldot_embed_text_sum(const int64_t* d_ids, const int64_t* d_pos_ids, int64_t pos_batch_stride, const void* d_word,
                        const void* d_pos, const void* d_type0, float* d_out, int32_t B, int32_t L, int32_t H,
                        int32_t vocab, int32_t max_pos, int32_t dtype, void* stream);
int ldot_embed_scatter(const float* d_dx, const int64_t* d_ids, const int64_t* d_pos_ids, int64_t pos_batch_stride,
                       float* d_dword, float* d_dpos, int32_t B, int32_t L, int32_t H, int32_t vocab, int32_t max_pos,
                       void* stream);
int ldot_embed_image_pre(const float* d_lin, const float* d_box, const float* d_img_g, const float* d_img_b,
                         const float* d_pos_w, const float* d_pos_bias, const float* d_pos_g, const float* d_pos_b,
                         const float* d_type1, float* d_q, float* d_spre, int64_t rows, int32_t H, void* stream);
int ldot_pos_wgrad(const float* d_dq, const float* d_box, int64_t rows, int32_t H, float* d_dw, void* stream);
int ldot_inbatch_nll_bwd(const float* d_scores, const int64_t* d_pos, int64_t bq, int64_t bc, const float* d_upstream,
                         int32_t reduction, void* d_dscores, int64_t ld_ds, int32_t dtype, void* stream);
int ldot_sumsq(const float* d_g, int64_t n, float* d_out, void* stream);
int ldot_adamw(float* d_p, const float* d_g, float* d_m, float* d_v, void* d_p16, int64_t n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int32_t step, const float* d_sumsq, float max_norm,
               int32_t dtype, void* stream);

/* ---- fused forms of the training step and CUDA-graph support (ABI 6) ---------------------------------------------
 * ldot_linear_dropout   BertSelfOutput / BertOutput in training mode up to the LayerNorm (layer.py:111-115,152-156):
 *                       d_out = dropout(d_a . d_w^T + d_bias) (+ d_residual), the mask of ldot_dropout(site) applied in
 *                       the GEMM epilogue (element index row * N + col); 16-bit output, N % 8 == 0
 * ldot_linear_gelu_grad BertIntermediate in training mode (layer.py:133-136): with z = d_a . d_w^T + d_bias rounded to 16 bit,
 *                       d_out = GELU(z) and d_gp = GELU'(z), both 16-bit, from one kernel and one exponential: the next
 *                       GEMM needs GELU(z), backward needs only GELU'(z) (ldot_gemm epi 4 multiplies by it)
 * ldot_gelu_grad        the same pair from an elementwise pass over a stored pre-activation: d_out = GELU(z), and d_z_gp
 *                       (z on entry) holds GELU'(z) on exit.  What the training forward uses: at K = 768 the GEMM epilogue
 *                       that evaluates both is issue-bound (151 us against 60 + 51 us for GEMM + this pass, 17.6 k tokens)
 * ldot_layernorm_bwd_dropout   ldot_layernorm_bwd (16-bit dy / x / dx) for a LayerNorm whose input was
 *                       dropout(dense) + residual: d_dx gets the residual-branch gradient, d_dx_masked = mask * dx / keep
 *                       the dense-branch gradient (row pitch ld_dx), and d_dxsum sums the MASKED values (the dense bias
 *                       gradient) - LayerNorm backward + ldot_dropout + ldot_colsum16 in one pass
 * ldot_dropout_epoch    registers a DEVICE word that every later dropout-bearing launch mixes into its mask key when the
 *                       kernel RUNS (NULL: none).  A captured CUDA graph replays launch arguments of capture time; bumping
 *                       the word between replays gives every replay fresh masks while forward and backward of one replay
 *                       agree.  Process-wide setting; returns the previous pointer through the return value's absence:
 *                       call with NULL to clear.
 * ldot_adamw_dev        ldot_adamw whose step-dependent scalars come from device memory, d_hyper = { lr, 1 - beta1^t,
 *                       sqrt(1 - beta2^t) } (fp32), so that a captured step follows the learning-rate schedule         */
int ldot_linear_dropout(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias,
                        const void* d_residual, int64_t ldr, void* d_out, int64_t ldo, int64_t M, int32_t N, int32_t K,
                        int32_t dtype, float drop_p, uint64_t seed, int32_t site, void* stream);
int ldot_linear_gelu_grad(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, const float* d_bias, void* d_gp,
                          int64_t ld_gp, void* d_out, int64_t ldo, int64_t M, int32_t N, int32_t K, int32_t dtype,
                         void* stream);
int ldot_gelu_grad(void* d_z_gp, void* d_out, int64_t n, int32_t dtype, void* stream);
int ldot_layernorm_bwd_dropout(const void* d_dy, int64_t ld_dy, const void* d_x, int64_t ld_x, const float* d_gamma,
                               void* d_dx, void* d_dx_masked, int64_t ld_dx, float* d_dgamma, float* d_dbeta,
                               float* d_dxsum, int64_t rows, int32_t H, float drop_p, uint64_t seed, int32_t site,
                               int32_t dtype, void* stream);
int ldot_dropout_epoch(const uint32_t* d_epoch);
int ldot_adamw_dev(float* d_p, const float* d_g, float* d_m, float* d_v, void* d_p16, int64_t n, const float* d_hyper,
                   float beta1, float beta2, float eps, float weight_decay, const float* d_sumsq, float max_norm,
                   int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDOT_H_ */
