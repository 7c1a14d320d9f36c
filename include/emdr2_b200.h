/*
 * emdr2_b200 — C ABI of the B200-native retrieve-and-read hot path.
 *
 * Plain C, no torch types: device/host pointers, sizes, an opaque handle and a cudaStream_t passed
 * as void*.  Every entry point returns 0 on success or a negative EMDR2_E* code; the message of the
 * last failure on the calling thread is available from emdr2_last_error().  Nothing throws across
 * the ABI and nothing calls exit().  All GPU work is enqueued on the caller's stream; the library
 * never calls cudaDeviceSynchronize() and never copies between host and device except in the
 * *_host convenience entry point that says so.
 *
 * What each entry point replaces in the reference (DevSinghSachan/emdr2 @ edb8cf67):
 *
 *   emdr2_mips_create / destroy   DistributedBruteForceIndex.__init__ / reset_index
 *                                 (megatron/data/emdr2_index.py:200-230) and
 *                                 FaissMIPSIndex._set_mips_index (:111-140)
 *   emdr2_mips_set_shard          DistributedBruteForceIndex.add_embed_data (:241-266): one row
 *                                 range of the torch.chunk split (:252) resident on one GPU, plus
 *                                 the row -> doc-id map (:258-260)
 *   emdr2_mips_search             DistributedBruteForceIndex.search_mips_index (:268-305) for one
 *                                 shard: Q·Eᵀ (:281) + top-k (:295) + id mapping (:298-303), and
 *                                 FaissMIPSIndex.search_mips_index (:182-197, IndexFlatIP.search)
 *   emdr2_mips_merge              the gather of per-GPU score slabs into C[nq,N] followed by the
 *                                 global torch.topk (:284-295); here a k-way merge of per-shard
 *                                 top-k lists (the payload of one all-gather)
 *   emdr2_mips_search_host        FaissMIPSIndex.search_mips_index's host round trip (:188,196):
 *                                 numpy queries in, numpy (distances, ids) out
 *
 * Ranking contract (the reference leaves tie order undefined — torch.topk on fp16 scores, FAISS heap
 * order): score[q,i] = sum_j Q[q,j]*E[i,j] accumulated in fp32 on the tensor cores; results are the
 * first k rows under (score descending, then row ascending inside a shard / id ascending across
 * merged lists).  A caller that stores each shard's rows in ascending id order therefore gets a
 * global (score desc, id asc) ranking for any number of shards.  Rows whose score is NaN are never
 * returned.  If fewer than k rows exist the tail is filled with score = -inf, id = -1 (FAISS
 * convention).
 */
#ifndef EMDR2_B200_H_
#define EMDR2_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EMDR2_API __attribute__((visibility("default")))
#else
#define EMDR2_API
#endif

#define EMDR2_OK 0
#define EMDR2_EINVAL (-1)   /* bad argument (null pointer, unsupported shape, misalignment) */
#define EMDR2_ECUDA (-2)    /* a CUDA runtime/driver call failed */
#define EMDR2_ENOMEM (-3)   /* host or device allocation failed */
#define EMDR2_ESTATE (-4)   /* call sequence error (e.g. search before set_shard) */
#define EMDR2_EUNSUPPORTED (-5) /* not an sm_100 device / feature not built */

#define EMDR2_DTYPE_FP16 0
#define EMDR2_DTYPE_BF16 1

/* Limits of the fused scan kernel (one launch). Larger nq is served by looping inside
 * emdr2_mips_search; larger k is rejected with EMDR2_EINVAL. */
#define EMDR2_MIPS_MAX_K 64
#define EMDR2_MIPS_QUERIES_PER_PASS 64

/* Message of the last error raised on the calling thread ("" if none). Never NULL. */
EMDR2_API const char* emdr2_last_error(void);

/* Library build information: "emdr2_b200 <version> sm_100a nvcc <x.y>". */
EMDR2_API const char* emdr2_version(void);

/* Create a search handle for embeddings of dimension d (d % 8 == 0, 8 <= d <= 1024) stored as
 * dtype (EMDR2_DTYPE_*) on CUDA device `device`.  Allocates the handle's small device workspace. */
EMDR2_API int emdr2_mips_create(int d, int dtype, int device, void** out_handle);

/* Bind one evidence shard: dev_rows is [n, d] row-major, 16-byte aligned, caller-owned device
 * memory that must stay valid until the next set_shard/destroy.  dev_ids is [n] int64 doc ids on
 * the device, or NULL meaning id = id_base + row.  n may be 0 (empty shard). */
EMDR2_API int emdr2_mips_set_shard(void* handle, const void* dev_rows, const int64_t* dev_ids, int64_t n,
                         int64_t id_base);

/* Top-k inner-product search of dev_q [nq, d] (same dtype as the shard, device memory) against the
 * bound shard.  Writes dev_scores [nq, k] fp32 and dev_ids [nq, k] int64, sorted best first.
 * Asynchronous on `cuda_stream` (a cudaStream_t; NULL = default stream). 1 <= k <= EMDR2_MIPS_MAX_K. */
EMDR2_API int emdr2_mips_search(void* handle, const void* dev_q, int nq, int k, float* dev_scores,
                      int64_t* dev_ids, void* cuda_stream);

/* Same search with HOST buffers: copies host_q to the device, searches, copies the results back and
 * synchronises the stream before returning (the FAISS-style numpy round trip). */
EMDR2_API int emdr2_mips_search_host(void* handle, const void* host_q, int nq, int k, float* host_scores,
                           int64_t* host_ids, void* cuda_stream);

/* k-way merge of `parts` sorted-or-unsorted top-k lists: scores [parts, nq, k] fp32 and
 * ids [parts, nq, k] int64 in device memory (entries with id < 0 are padding) into
 * out_scores/out_ids [nq, k] ranked (score desc, id asc).  Handle-free; asynchronous on the stream. */
EMDR2_API int emdr2_mips_merge(const float* dev_scores, const int64_t* dev_ids, int parts, int nq, int k,
                     float* dev_out_scores, int64_t* dev_out_ids, void* cuda_stream);

/* Tuning / introspection. Options: "probe" (0/1, seed thresholds from a first probing tile),
 * "share" (0/1, cross-CTA threshold sharing), "max_ctas" (0 = all SMs), "stats" (0/1, count
 * appends/compactions), "timing" (0/1, bracket every scan launch with CUDA events on the stream).
 * Stats (of the last search, valid after the stream has been synchronised): "ctas", "tiles",
 * "stages", "smem_bytes", "sm_count", "appends", "compactions", "probe_wait_ns"; with "timing" on:
 * "scan_launches" and "scan_ns" (sum of scan-kernel durations since switched on; reading resets). */
EMDR2_API int emdr2_mips_set_option(void* handle, const char* name, int64_t value);
EMDR2_API int emdr2_mips_get_stat(void* handle, const char* name, int64_t* out_value);

EMDR2_API int emdr2_mips_destroy(void* handle);

/* ------------------------------------------------------------------------------------------------
 * Transformer-block operators (BERT towers and the T5 reader).  All tensors are 16-bit (dtype =
 * EMDR2_DTYPE_*), row-major, device memory, 16-byte aligned; all calls are asynchronous on
 * `cuda_stream` and run on the calling thread's current CUDA device.
 * ---------------------------------------------------------------------------------------------- */

#define EMDR2_GEMM_BIAS 1      /* + bias[n]                                                        */
#define EMDR2_GEMM_GELU 2      /* exact (erf) GeLU after the bias                                   */
#define EMDR2_GEMM_RESIDUAL 4  /* + residual[m, n] after the activation                             */

/* d[m,n] = epilogue(a[m,k] . b[n,k]^T): torch.nn.functional.linear with fp32 accumulation and the
 * bias / GeLU / residual-add fused into the store.  Replaces mpu.ColumnParallelLinear /
 * RowParallelLinear at model-parallel size 1 (megatron/mpu/layers.py:170-363) with the bias-GeLU
 * (megatron/model/transformer.py:99-104) and bias-dropout-add at p=0 (:397-419) of the reference's
 * layer, and parallel_lm_logits (megatron/model/language_model.py:28-42).  lda/ldb/ldd/ldr are row
 * pitches in elements (multiples of 8); n a multiple of 8. */
EMDR2_API int emdr2_gemm(int dtype, const void* a, int64_t lda, const void* b, int64_t ldb, void* d,
                         int64_t ldd, const void* bias, const void* residual, int64_t ldr, int m,
                         int n, int k, int flags, void* cuda_stream);

#define EMDR2_GEMM_ACCUM_F32 8  /* d is fp32 [m, ldd] and the product is atomically ADDED to it      */
#define EMDR2_GEMM_GELU_BWD 16  /* d = acc * GeLU'(aux[m, n])  (aux = saved pre-activation)          */
#define EMDR2_GEMM_PREACT 32    /* also store acc + bias (before GeLU) to preact[m, n]               */

/* General form used by the training path.  a_mn / b_mn != 0: that operand is stored with the
 * contraction index as its ROW index (a as [k, m], b as [k, n], row-major), so the backward products
 * read the forward tensors in place:
 *   dX[m,k] = dY[m,n] . W[n,k]      -> emdr2_gemm_ex(a = dY, b = W  with b_mn = 1)
 *   dW[n,k] += dY[m,n]^T . X[m,k]   -> emdr2_gemm_ex(a = dY with a_mn = 1, b = X with b_mn = 1,
 *                                      EMDR2_GEMM_ACCUM_F32, splits > 1: split-K over the tokens)
 * (autograd of mpu.ColumnParallelLinear / RowParallelLinear, megatron/mpu/layers.py:170-363).
 * aux: residual (EMDR2_GEMM_RESIDUAL) or saved pre-activation (EMDR2_GEMM_GELU_BWD, the backward of
 * transformer.py:99-104 fused into the dA = dY . W2 product). */
EMDR2_API int emdr2_gemm_ex(int dtype, const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb,
                            int b_mn, void* d, int64_t ldd, const void* bias, const void* aux,
                            int64_t ld_aux, void* preact, int64_t ld_preact, int m, int n, int k,
                            int flags, int splits, void* cuda_stream);

/* Fused attention forward, head dimension 64:
 *   o[b,i,h,:] = softmax_j(mask(scale * q[b,i,h,:].k[b,j,h,:])) . v[b,j,h,:]
 * q/o are [batch*sq, >= heads*64] and k/v [batch*sk, >= heads*64] row-major views (row pitches
 * ldq/ldk/ldv/ldo in elements, head h in columns h*64 .. h*64+63) - e.g. column blocks of a fused
 * QKV projection.  Mask = q_pad[b,i] | k_pad[b,j] | (causal && j > i) (byte vectors, 1 = masked,
 * NULL = none); a masked score is replaced by -10000 exactly as the reference's
 * attention_mask_func does (megatron/model/bert_model.py:31-33, t5_model.py:28-30), then softmax.
 * Replaces the baddbmm -> scale-mask-softmax -> bmm core of ParallelAttention.forward
 * (megatron/model/transformer.py:301-383) with attention dropout off.  lse (optional) receives
 * log(sum_j exp(masked score)) as [batch, heads, sq] fp32.
 * q_live [batch, ceil(sq/128)] / k_live [batch, ceil(sk/128)] (optional bytes): 0 marks a block of
 * 128 queries / keys that is all padding; such key blocks are skipped (their probabilities are
 * exactly 0 for every non-padding query) and such query blocks store zeros; every batch entry needs
 * at least one live key block.  Outputs at non-padding queries are unchanged; padding rows (which
 * no consumer reads) are no longer the reference's uniform average.  NULL = exact reference
 * behaviour everywhere. */
EMDR2_API int emdr2_attention_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                  const void* v, int64_t ldv, void* o, int64_t ldo, int batch,
                                  int heads, int sq, int sk, const uint8_t* q_pad,
                                  const uint8_t* k_pad, const uint8_t* q_live, const uint8_t* k_live,
                                  int causal, float scale, float* lse, void* cuda_stream);

/* y[r,:] = LayerNorm(x[r,:]) * gamma + beta with fp32 statistics (biased variance), h % 8 == 0,
 * h <= 1024: mpu.LayerNorm (megatron/mpu/layers.py:28-36).  mean/rstd: optional [rows] fp32. */
EMDR2_API int emdr2_layernorm_fwd(int dtype, const void* x, int64_t ldx, const void* gamma,
                                  const void* beta, void* y, int64_t ldy, int rows, int h, float eps,
                                  float* mean, float* rstd, void* cuda_stream);

/* out[t,:] = word[ids[t]] + pos[t % seq] (+ type_emb[types[t]]): Embedding.forward with dropout off
 * (megatron/model/language_model.py:169-181).  ids/types: int64 [tokens] device arrays. */
EMDR2_API int emdr2_embedding_fwd(int dtype, const int64_t* ids, const int64_t* types,
                                  const void* word, const void* pos, const void* type_emb, void* out,
                                  int tokens, int seq, int h, int vocab, int num_types,
                                  void* cuda_stream);

/* logprob[r] = logits[r, labels[r]] - logsumexp(logits[r, :]) (and lse[r], optional) in one pass:
 * the log_softmax + gather of get_loss_and_retriever_utility / get_kl_div_retriever
 * (tasks/openqa/e2eqa/train_e2eqa.py:82-98,196-208) and of the reader's CrossEntropyLoss (:156-160).
 * logits [rows, vocab] 16-bit with row pitch ld; labels int64 [rows] (out of range -> logprob 0). */
EMDR2_API int emdr2_token_logprob(int dtype, const void* logits, int64_t ld, const int64_t* labels,
                                  float* logprob, float* lse, int rows, int vocab, void* cuda_stream);

/* ---- backward of the block operators (training path).  Activation gradients are 16-bit; parameter
 * gradients are ACCUMULATED (atomic adds) into caller-owned fp32 buffers, Megatron main-grad style. */

/* dq, dk, dv from dout and the forward's q, k, v, o, lse (autograd of transformer.py:301-383).
 * Layouts, masks, live maps and scale as in emdr2_attention_fwd; dvec_ws: [batch, heads, sq] fp32
 * scratch.  No gradient flows into masked scores (masked_fill semantics). */
EMDR2_API int emdr2_attention_bwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                  const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout,
                                  int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                                  int64_t lddv, int batch, int heads, int sq, int sk, const uint8_t* q_pad,
                                  const uint8_t* k_pad, const uint8_t* q_live, const uint8_t* k_live,
                                  int causal, float scale, const float* lse, float* dvec_ws,
                                  void* cuda_stream);

/* LayerNorm backward: dx = LN'(dy) (+ dres, the gradient on the residual branch); dgamma += ...,
 * dbeta += ... (fp32, may be NULL).  mean/rstd from emdr2_layernorm_fwd. */
EMDR2_API int emdr2_layernorm_bwd(int dtype, const void* dy, int64_t ldy, const void* x, int64_t ldx,
                                  const void* gamma, const float* mean, const float* rstd, const void* dres,
                                  int64_t ldr, void* dx, int64_t lddx, float* dgamma, float* dbeta, int rows,
                                  int h, void* cuda_stream);

/* out[n] += sum_m dy[m, n]: bias gradients. */
EMDR2_API int emdr2_colsum(int dtype, const void* dy, int64_t ld, float* out, int rows, int n, void* cuda_stream);

/* dlogits[r, :] = g[r] * (onehot(labels[r]) - softmax(logits[r, :])): backward of
 * emdr2_token_logprob given g = dLoss/dlogprob and the saved lse. */
EMDR2_API int emdr2_token_logprob_bwd(int dtype, const void* logits, int64_t ld, const int64_t* labels,
                                      const float* lse, const float* g, void* dlogits, int64_t ldd, int rows,
                                      int vocab, void* cuda_stream);

/* dword[ids[t]] += dx[t]; dpos[t % seq] += dx[t]; dtype_emb[types[t]] += dx[t] (fp32, any may be NULL). */
EMDR2_API int emdr2_embedding_bwd(int dtype, const void* dx, const int64_t* ids, const int64_t* types,
                                  float* dword, float* dpos, float* dtype_emb, int tokens, int seq, int h,
                                  int vocab, int num_types, void* cuda_stream);

/* Variable-length attention forward over TOKEN-PACKED activations (csrc/attention_varlen.cu): q [q_rows, >= heads*64],
 * k / v [k_rows, >= heads*64] and o [o_rows, >= heads*64] are matrices in which sequences follow each other without
 * padding; `items` (device, n_items x 8 int32: q_row0, q_valid (1..128), k_row0, k_len (>= 1), head, o_row0, lse_idx0,
 * reserved) lists the (128-query tile, head, key range) products to compute — the self-attention of every sequence of
 * a batch in ONE launch (transformer.py:301-383 over emdr2_model.py:118-120,148-149's padded rectangles), or the
 * key ranges of a FiD cross-attention, to be merged by the caller with the lse weights.  No masks: a key outside the
 * item's range has probability 0 (what masked_fill(-10000) + softmax gives a padding key).  Rows of a partial tile
 * beyond q_valid are not written.  lse (optional): natural-log sum-exp per query at lse_idx0 + row. */
EMDR2_API int emdr2_attention_varlen_fwd(int dtype, const void* q, int64_t ldq, int64_t q_rows, const void* k,
                                         int64_t ldk, const void* v, int64_t ldv, int64_t k_rows, void* o, int64_t ldo,
                                         int64_t o_rows, int heads, const int32_t* dev_items, int n_items, float scale,
                                         float* lse, void* cuda_stream);

/* emdr2_embedding_fwd with explicit positions: pos_ids int32 [tokens] (NULL = t % seq), for packed sequences. */
EMDR2_API int emdr2_embedding_fwd_pos(int dtype, const int64_t* ids, const int64_t* types, const void* word,
                                      const void* pos, const void* type_emb, void* out, int tokens, int seq, int h,
                                      int vocab, int num_types, const int32_t* pos_ids, int max_pos,
                                      void* cuda_stream);

/* ---- dropout (csrc/dropout.cuh).  The reference trains with torch dropout, p = 0.1, on the attention
 * probabilities (megatron/model/transformer.py:345-346), on every bias-add-residual (transformer.py:397-419,
 * 511-515) and on the embedding sum (language_model.py:181), and stores the masks for the backward pass.
 * Here the mask is a counter-based function of (seed, offset, row, column): the caller picks one `offset`
 * per dropout call site and training step and hands the SAME (p, seed, offset) to the backward entry point,
 * which regenerates the mask.  p_eff = round(p * 2^32) / 2^32; kept values are scaled by 1 / (1 - p_eff).
 * `colhash` is a caller-owned device table of column hashes for `seed` (emdr2_dropout_colhash), at least
 * as long as the op's column count rounded up to 128. */
EMDR2_API int emdr2_dropout_colhash(uint64_t seed, uint32_t* dev_table, int n, void* cuda_stream);

/* mask[r, c] = 1 where (r, c) is kept, [rows, cols] bytes: for tests that replay a mask in a reference. */
EMDR2_API int emdr2_dropout_mask(float p, uint64_t seed, uint64_t offset, const uint32_t* colhash,
                                 uint8_t* dev_mask, int64_t rows, int cols, void* cuda_stream);

/* out = residual + dropout(y) over [rows, cols] 16-bit tensors (residual may be NULL, out may alias y):
 * bias_dropout_add once the bias is in y; with y = the incoming gradient and residual = NULL, its backward. */
EMDR2_API int emdr2_dropout_add(int dtype, const void* y, int64_t ldy, const void* residual, int64_t ldr,
                                void* out, int64_t ldo, int rows, int cols, float p, uint64_t seed,
                                uint64_t offset, const uint32_t* colhash, void* cuda_stream);

/* emdr2_attention_fwd / emdr2_attention_bwd with attention dropout: P is multiplied by keep / (1 - p) after
 * the softmax normalisation (lse is that of the undropped probabilities); dropout plane row =
 * (b * heads + head) * sq + query, column = key.  p = 0 is exactly the entry points above. */
EMDR2_API int emdr2_attention_fwd_dropout(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                          const void* v, int64_t ldv, void* o, int64_t ldo, int batch, int heads,
                                          int sq, int sk, const uint8_t* q_pad, const uint8_t* k_pad,
                                          const uint8_t* q_live, const uint8_t* k_live, int causal, float scale,
                                          float* lse, float p, uint64_t seed, uint64_t offset,
                                          const uint32_t* colhash, void* cuda_stream);
EMDR2_API int emdr2_attention_bwd_dropout(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                          const void* v, int64_t ldv, const void* o, int64_t ldo, const void* dout,
                                          int64_t lddo, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                                          int64_t lddv, int batch, int heads, int sq, int sk, const uint8_t* q_pad,
                                          const uint8_t* k_pad, const uint8_t* q_live, const uint8_t* k_live,
                                          int causal, float scale, const float* lse, float* dvec_ws, float p,
                                          uint64_t seed, uint64_t offset, const uint32_t* colhash,
                                          void* cuda_stream);

/* ------------------------------------------------------------------------------------------------
 * Host-side integer work of the step (no device code, callable without a GPU).
 * ---------------------------------------------------------------------------------------------- */

/* Passage -> model-input formatting for one batch: replaces the B*K Python iterations of
 * EMDR2Model.postprocess (megatron/model/emdr2_model.py:250-303) and the list builders it calls -
 * context_bert_format (megatron/data/orqa_wiki_dataset.py:86-120), query_extended_context_t5_format
 * (emdr2_model.py:306-361) and query_single_context_t5_format (:364-376) - element for element.
 *   query_uid [bsz], query_ids [bsz, query_stride] with query_len [bsz] valid tokens per row;
 *   candidates of question b are cand_begin[b] .. cand_begin[b+1]-1 (k or k+1 of them, best first);
 *   cand_id [n_cand] evidence ids (a candidate whose id equals the question's uid is dropped, as are
 *   candidates beyond the first k_keep kept ones); cand_meta [n_cand, 6] = title_len, n_docs (1..3:
 *   the passage and its neighbours, inverted_title_index.py:22-37), main_idx (0, 1 or -1: which of
 *   them is the passage), doc_len[3]; tokens = every candidate's title followed by its docs, back to
 *   back (n_tokens in total).
 * Outputs (host memory, caller-owned, fully overwritten): ctx_ids / ctx_types [bsz, k_keep, seq_ret],
 * extended / single [bsz*k_keep, seq]; max_len[3] (optional) = longest non-padding prefix in ctx_ids,
 * extended, single; row_len [3, bsz*k_keep] (optional) = that prefix length for every row of the
 * three layouts (what lets the towers run length-bucketed without a device sync).  EMDR2_EINVAL when a question keeps fewer than k_keep candidates or
 * question + title do not fit in seq (the reference would build a ragged tensor / overflow there). */
EMDR2_API int emdr2_format_passages(int32_t bsz, int32_t k_keep, const int64_t* query_uid,
                                    const int64_t* query_ids, int64_t query_stride,
                                    const int64_t* query_len, const int32_t* cand_begin,
                                    const int64_t* cand_id, const int32_t* cand_meta,
                                    const int64_t* tokens, int64_t n_tokens, int32_t seq_ret,
                                    int32_t seq, int64_t cls_id, int64_t sep_id, int64_t pad_id,
                                    int64_t* ctx_ids, int64_t* ctx_types, int64_t* extended,
                                    int64_t* single, int32_t* max_len, int32_t* row_len);

/* Same formatting with the tokens read IN PLACE from flat token stores (what the reference's memory-mapped
 * indexed datasets are: one token buffer + per-document offsets, megatron/data/indexed_dataset.py): no
 * per-passage array objects, no concatenation.  piece_offset [n_cand, 4] = element offset of the
 * candidate's title in title_tokens and of its up to three passages in doc_tokens (lengths in cand_meta);
 * token_bytes = 2 (uint16), 4 (int32) or 8 (int64) - the element type of both stores. */
EMDR2_API int emdr2_format_passages_flat(int32_t bsz, int32_t k_keep, const int64_t* query_uid,
                                         const int64_t* query_ids, int64_t query_stride,
                                         const int64_t* query_len, const int32_t* cand_begin,
                                         const int64_t* cand_id, const int32_t* cand_meta,
                                         const int64_t* piece_offset, const void* title_tokens,
                                         int64_t n_title_tokens, const void* doc_tokens,
                                         int64_t n_doc_tokens, int32_t token_bytes, int32_t seq_ret,
                                         int32_t seq, int64_t cls_id, int64_t sep_id, int64_t pad_id,
                                         int64_t* ctx_ids, int64_t* ctx_types, int64_t* extended,
                                         int64_t* single, int32_t* max_len, int32_t* row_len);

/* Process-wide switches of the block operators.  "gemm_pair" (initial value from the environment
 * variable EMDR2_GEMM_PAIR): 1 = large K-major products with a 16-bit output run on CTA pairs
 * (tcgen05.mma.cta_group::2, 256 x 256 tiles, residual box in its own staging buffer) instead of one
 * CTA per 128 x 256 tile; 2 (default) = only where that is measured to win (residual / aux epilogues
 * over >= 100 k rows); 0 = never.  Results are bit-identical between the two kernels (same products, same fp32
 * accumulation order per element). */
/* "gemm_wide" (EMDR2_GEMM_WIDE): sixteen instead of eight epilogue warps in the one-CTA kernel for 16-bit outputs —
 * 1 (default) = where the epilogue has arithmetic or an aux tile's latency to hide (GeLU, pre-activation side output,
 * GeLU backward), 2 = whenever eligible (no fp32 accumulation, no split-K), 0 = never.  Bit-identical results. */
/* "gemm_max_ctas": > 0 caps the persistent GEMM grids at that many CTAs (0 = one per SM): leave SMs to a collective
 * kernel that overlaps with the backward pass instead of running the grid's last CTAs as a second wave.  The cap also
 * decides how well the weight-gradient products fill the machine (output tiles x split-K factor vs CTAs): with the
 * step's 18 / 54 / 72-tile products 144 CTAs (4 SMs for NCCL, NCCL_MAX_CTAS=4) are filled exactly, 140 are not. */
EMDR2_API int emdr2_ops_set_option(const char* name, int64_t value);
EMDR2_API int emdr2_ops_get_option(const char* name, int64_t* out_value);

/* Measurement aid: with timing enabled every launch of the block operators made by this
 * process is bracketed by CUDA events on its stream (backward passes run on autograd worker threads).  emdr2_ops_timing_read sums the launch
 * durations (ns), launch count and algorithmic FLOPs of one kernel kind since the last read and
 * restarts the accumulation (blocks until the last timed launch has finished). */
#define EMDR2_KIND_GEMM 0
#define EMDR2_KIND_ATTENTION 1
#define EMDR2_KIND_ROWOP 2
#define EMDR2_KIND_COUNT 3
EMDR2_API int emdr2_ops_timing(int enable);
/* Credit algorithmic flops to a kind whose entry point cannot know them (emdr2_attention_varlen_fwd). */
EMDR2_API int emdr2_ops_timing_add_flops(int kind, double flops);
EMDR2_API int emdr2_ops_timing_read(int kind, int64_t* out_ns, int64_t* out_launches, double* out_flops);

#ifdef __cplusplus
}
#endif
#endif /* EMDR2_B200_H_ */
