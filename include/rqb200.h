/* librqb200 -- C ABI of the B200-native RQ-VAE residual-quantisation hot path.
 *
 * The reference (EdoardoBotta/RQ-VAE-Recommender) has no FFI layer: its boundary is the Python module API
 * (modules/quantize.py, modules/rqvae.py, init/kmeans.py, distributions/gumbel.py, modules/encoder.py).
 * This header is the C boundary those replacement modules bind (via ctypes, see
 * rq_vae_recommender_b200/_lib.py and INTEGRATION.md).  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host (`const float* const* codebooks` is a HOST
 *     array of L device pointers); row-major storage; strides/leading dimensions are in ELEMENTS;
 *   - no allocation, no ownership transfer: the caller (PyTorch) allocates outputs and workspaces
 *     (sizes from the *_workspace_bytes queries) and keeps them alive until the stream work completes;
 *   - `stream` is a cudaStream_t; all work is enqueued on it, nothing synchronises the host;
 *   - every function returns 0 on success or an RQB_ERR_* code; rqb200_last_error() gives the text
 *     (thread-local).  No exceptions cross the boundary.  No global mutable state besides that string.
 */
#ifndef RQB200_H
#define RQB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RQB_OK 0
#define RQB_ERR_INVALID 1
#define RQB_ERR_CUDA 2
#define RQB_ERR_UNSUPPORTED 3
#define RQB_ERR_WORKSPACE 4

#define RQB_MAX_LEVELS 8

/* forward modes: numbering of modules/quantize.py:16-20 (QuantizeForwardMode), 0 = eval-mode lookup */
#define RQB_MODE_EVAL 0
#define RQB_MODE_GUMBEL 1
#define RQB_MODE_STE 2
#define RQB_MODE_ROTATION 3

int rqb200_version(void);
const char* rqb200_last_error(void);
int rqb200_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- fused L-level residual quantisation, exact fp32 -------------------------------------------------
 * Replaces L x Quantize.forward (modules/quantize.py:104-163: distance :113-117, argmin :128, eval lookup
 * :159-161, STE :137-139, rotation trick :140-153 + :34-50, QuantizeLoss modules/loss.py:38-41) chained by
 * RqVae.get_semantic_ids (modules/rqvae.py:122-139: residual update :130, loss accumulation :128).
 * x is the encoder output [B,D] (ldx >= D).  All outputs are optional (NULL = not produced):
 *   ids [B,L] int64; embeddings / residuals [L,B,D] (the Python side returns [B,D,L] views);
 *   emb_sum [B,D] = sum_l emb_out (rqvae.py:146); emb_norms [B,L] = ||emb_out|| (rqvae.py:158); loss [B].
 * Tokenisation (modules/tokenizer/semids.py:125) is mode = EVAL with only `ids`. */
size_t rqb200_rq_workspace_bytes(int D, int K, int L);
int rqb200_rq_forward(int mode, const float* x, int64_t ldx, const float* const* codebooks, int B, int D, int K,
                      int L, float beta, int64_t* ids, float* embeddings, float* residuals, float* emb_sum,
                      float* emb_norms, float* loss, void* workspace, size_t workspace_bytes, void* stream);

/* The same outputs from GIVEN ids (no distance computation): the streaming half of the chain.  With the ids of the
 * tensor-core tokeniser (identical to the exact kernel's) this is how a large training-mode batch runs: tokenise + this pass
 * instead of CUDA-core distances; outputs are bit-identical to rqb200_rq_forward's for the same ids. */
int rqb200_rq_forward_from_ids(int mode, const float* x, int64_t ldx, const float* const* codebooks, const int64_t* ids,
                               int B, int D, int K, int L, float beta, float* embeddings, float* residuals, float* emb_sum,
                               float* emb_norms, float* loss, void* stream);

/* Backward of the above (autograd of quantize.py:130-161 + loss.py:38-41 through the chain rqvae.py:125-132;
 * formulas SURVEY appendix A.3).  Upstream grads come with element strides so expanded / permuted torch
 * tensors need no copy: g_emb[b,d,l] = g_emb[b*ge_sB + d*ge_sD + l*ge_sL] (NULL = zero), same for g_res;
 * g_loss[b*gl_sB].  g_x [B,D] is written; g_codebooks[l] [K,D] are ACCUMULATED (caller zero-fills). */
int rqb200_rq_backward(int mode, const float* x, int64_t ldx, const float* const* codebooks, const int64_t* ids,
                       int B, int D, int K, int L, float beta, const float* g_emb, int64_t ge_sB, int64_t ge_sD,
                       int64_t ge_sL, const float* g_res, int64_t gr_sB, int64_t gr_sD, int64_t gr_sL,
                       const float* g_loss, int64_t gl_sB, float* g_x, float* const* g_codebooks, void* stream);

/* ---- tensor-core tokeniser (tcgen05 candidate filter + exact fp32 re-rank), see csrc/rq_tc.cu --------
 * Same result contract as rqb200_rq_forward(mode=EVAL, ids only).  `prepare` converts the codebooks once
 * (fp16 copies, norms, inter-level Gram tables) into `state`; `run` consumes x [B,D] fp32. */
size_t rqb200_tokenize_tc_state_bytes(int D, int K, int L);
int rqb200_tokenize_tc_supported(int D, int K, int L);
int rqb200_tokenize_tc_prepare(const float* const* codebooks, int D, int K, int L, void* state, size_t state_bytes,
                               void* stream);
int rqb200_tokenize_tc_run(const float* x, int64_t ldx, int B, const void* state, int D, int K, int L,
                           int64_t* ids, int* stats, void* stream);

/* ---- k-means codebook initialisation (init/kmeans.py) ------------------------------------------------
 * assign_accumulate = one Lloyd assignment pass with the direct (x-c)^2 distance of kmeans.py:40-43 plus the
 * per-cluster sums (fp64) and counts that kmeans.py:48-58 derives with a Python loop; in the sharded setting
 * the caller all-reduces sums/counts between the two calls.  finalize writes the new centroids in place
 * (mean, or x[reseed_rows[k]] for an empty cluster, kmeans.py:50-54; reseed_rows may be NULL) and the
 * max centroid shift of kmeans.py:68 into *max_shift (device float). */
int rqb200_kmeans_assign_accumulate(const float* x, int64_t ldx, const float* centroids, int B, int D, int K,
                                    int64_t* assignment, double* sums, int* counts, void* workspace,
                                    size_t workspace_bytes, void* stream);
int rqb200_kmeans_finalize(const double* sums, const int* counts, const float* x, int64_t ldx,
                           const int64_t* reseed_rows, float* centroids, int K, int D, float* max_shift,
                           void* stream);

/* ---- dense fp32 helpers -------------------------------------------------------------------------------
 * sgemm: C = epi(alpha * op(A) op(B) + beta * C), row-major, op = transpose flag; relu and the (mask > 0)
 * epilogue fuse the ReLU forward / backward of modules/encoder.py:27-29.  Also the W@C and gradient GEMMs of
 * the Gumbel path (quantize.py:135). */
int rqb200_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int relu, const float* mask,
                 int64_t ldmask, void* stream);
int rqb200_row_sqnorm(const float* c, int K, int D, float* out, void* stream);
/* dots [B,K] (= x @ C^T) -> dist in place (quantize.py:113-117) + first-index argmin (quantize.py:128) */
int rqb200_dist_finish(float* dots, const float* x, int64_t ldx, const float* cc, int B, int D, int K, int64_t* ids,
                       void* stream);
/* W = softmax((-dist + G(U)) / T)  (distributions/gumbel.py:8-20 with U injected; quantize.py:132-134) */
int rqb200_gumbel_softmax_fwd(const float* dist, const float* uniform, float* weights, int B, int K,
                              float temperature, void* stream);
int rqb200_gumbel_row_finish(const float* x, int64_t ldx, const float* E, int B, int D, float beta, float* loss,
                             void* stream);
int rqb200_gumbel_bwd_ge(const float* g_out, int64_t go_sB, int64_t go_sD, const float* g_loss, int64_t gl_sB,
                         const float* x, int64_t ldx, const float* E, float* gE, int B, int D, void* stream);
int rqb200_gumbel_bwd_softmax(const float* weights, float* gw_inout, int B, int K, float temperature, float* rowsum,
                              float* colsum, void* stream);
int rqb200_gumbel_bwd_gx(float* acc_inout, const float* x, int64_t ldx, const float* E, const float* g_loss,
                         int64_t gl_sB, const float* rowsum, float beta, int B, int D, void* stream);
int rqb200_gumbel_bwd_gc(float* gC_inout, const float* C, const float* colsum, int K, int D, void* stream);
/* modules/normalize.py:6-7 (F.normalize p=2) and its backward */
int rqb200_l2norm_fwd(const float* x, float* y, float* norms, int B, int D, float eps, void* stream);
int rqb200_l2norm_bwd(const float* gy, const float* y, const float* norms, float* gx, int B, int D, float eps,
                      void* stream);

/* ---- bf16 tensor-core GEMM for the MLPs (modules/encoder.py:23-38), reduced-precision / AMP-like, forward only ---
 * The reference runs these Linears in bf16 when mixed precision is on (train_rqvae.py:36,69).  Operands live in HBM
 * as "images": [tile of 128 rows][64-wide k chunk][128 x 128 B], the exact K-major SWIZZLE_128B shared-memory layout
 * tcgen05 reads, so a pipeline stage is one contiguous TMA bulk copy.  K must be a multiple of 64.
 *   f32_to_bf16_image : fp32 row-major [rows,K] -> image (used for the input x AND for a weight W[N,K]);
 *   gemm_bf16         : Y = act(X W^T), fp32 accumulate; Y is written as the next layer's A image (N % 64 == 0)
 *                       and/or as fp32 row-major [M,N]. */
size_t rqb200_bf16_image_bytes(int rows, int K);
int rqb200_f32_to_bf16_image(const float* x, int64_t ldx, int rows, int K, void* image, void* stream);
int rqb200_gemm_bf16(const void* a_image, const void* w_image, int M, int N, int K, int relu, void* out_image,
                     float* out_f32, int64_t ldo, void* stream);

/* ---- split-precision tensor-core GEMM: fp32-accurate products on the fp16 tensor cores -----------------------------
 * The MLP Linears of modules/encoder.py:23-38 in their default (index-exact) precision and the two GEMMs of a
 * Gumbel-softmax level (modules/quantize.py:113-117,135).  Each operand row is scaled by a power of two and stored as
 * two fp16 images hi + lo (22 significant bits); C = act(A B^T) costs three tcgen05.mma per k-step (hi.hi + lo.hi +
 * hi.lo, fp32 accumulate) and is written as fp32 rows.
 *   split_image_bytes  : bytes of one operand buffer [hi image][lo image][row scales] for a [rows, K] matrix
 *   f32_to_split_image : fp32 [rows, K] (ld = ldx) -> buffer; transposed != 0 reads the operand as x[K, rows]
 *   gemm_split         : out[M, N] (ld = ldo) = act(A[M, K] . B[N, K]^T) from two such buffers; an optional mask[M, N]
 *                        zeroes the entries whose mask value is not > 0 (the ReLU' of a backward GEMM) */
size_t rqb200_split_image_bytes(int rows, int K);
int rqb200_f32_to_split_image(const float* x, int64_t ldx, int rows, int K, int transposed, void* image, void* stream);
int rqb200_gemm_split(const void* a_image, const void* b_image, int M, int N, int K, int relu, const float* mask,
                      int64_t ldm, float* out, int64_t ldo, void* stream);
/* split-K schedule for few output tiles and a long contraction (weight gradients, modules/encoder.py backward: M = out,
 * N = in, K = batch): gemm_split_k_slices returns the slice count that fills the SMs; workspace = slices * M * N floats
 * (unused when slices == 1); the partial sums are reduced in a fixed order. */
int rqb200_gemm_split_k_slices(int M, int N, int K);
int rqb200_gemm_split_k(const void* a_image, const void* b_image, int M, int N, int K, int slices, float* workspace,
                        float* out, int64_t ldo, void* stream);

/* ---- corpus id statistics (train_rqvae.py:279-289, modules/tokenizer/semids.py:94-108) ---------------- */
int rqb200_sid_histogram(const int64_t* ids, int B, int L, int K, int64_t* hist /* [L,K], zeroed here */,
                         void* stream);

/* Dedup column of the corpus table, modules/tokenizer/semids.py:94-108: rank[i] = number of rows j < i with the same id tuple
 * (the reference's O(N^2) compare), and in the same pass the diversity statistics of train_rqvae.py:276-283:
 * stats[0] = max rank (max_id_duplicates * N), stats[1] = number of distinct tuples, *entropy = -sum p log p over the distinct
 * tuples.  Direct-table algorithm: needs K^L <= 2^26 (workspace_bytes returns 0 otherwise and the call RQB_ERR_UNSUPPORTED:
 * the caller falls back to a sort).  ids [N,L] int64 row-major; ids outside [0,K) make a row its own group. */
size_t rqb200_sid_dedup_workspace_bytes(int N, int L, int K);
int rqb200_sid_dedup_rank(const int64_t* ids, int N, int L, int K, int64_t* rank /* [N] */, int* stats /* [2] */,
                          double* entropy /* [1] */, void* workspace, size_t workspace_bytes, void* stream);

/* Sequence tokenisation from the corpus table, semids.py:112-146 (`cached_ids[ids]`, -1 under the padding mask,
 * token_type_ids): out[b, s*C + c] = seq_mask[b,s] ? cached_ids[item_ids[b,s], c] : -1; token_type[b, s*C + c] = c.
 * seq_mask (bytes, 0 = padding) and token_type may be null; strides in elements. */
int rqb200_sid_gather(const int64_t* cached_ids, int64_t n_corpus, int C, const int64_t* item_ids, int64_t item_stride,
                      const unsigned char* seq_mask, int64_t mask_stride, int B, int S, int64_t* out /* [B, S*C] */,
                      int64_t* token_type /* [B, S*C] or null */, void* stream);

/* ---- constrained beam search, data side (modules/model.py:169-182 `_check_valid_prefix`, :340-376 the selection step) ----
 * sid_prefix_build  : corpus id table [N, C] -> one bitmap per prefix length (workspace layout: levels 1..C, 256-byte aligned
 *                     regions; sid_prefix_workspace_bytes returns 0 when K^C exceeds 2^33 bits)
 * sid_prefix_check  : valid[p] = some corpus row starts with prefix[p, :l]   (one bit test instead of the reference's
 *                     O(P N l) compare)
 * sid_beam_select   : one launch per hierarchy level h: score kp x nc candidate extensions per batch row (sampled token
 *                     log-probability + parent beam log-probability, -inf when the extended prefix is not in the corpus), keep
 *                     the k best in descending order, gather their ids into out_generated [B, k, h + 1] and return the parent
 *                     beam's global index b * kp + beam (the key/value-cache reorder index). */
size_t rqb200_sid_prefix_workspace_bytes(int C, int K);
int rqb200_sid_prefix_build(const int64_t* cached_ids, int64_t N, int C, int K, void* workspace, size_t ws_bytes, void* stream);
int rqb200_sid_prefix_check(const int64_t* prefix, int64_t row_stride, int64_t P, int l, int C, int K, const void* workspace,
                            unsigned char* valid, void* stream);
int rqb200_sid_beam_select(const int64_t* samples, const float* samp_log_p, const int64_t* generated, const float* log_probas,
                           int B, int kp, int nc, int h, int k, int C, int K, const void* prefix_workspace,
                           int64_t* out_generated, float* out_log_probas, int64_t* out_parent, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RQB200_H */
