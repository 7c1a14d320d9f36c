// Row-wise HBM-bound operators of the transformer blocks: LayerNorm and the embedding sum.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emdr2 {

// y[r,:] = (x[r,:] - mean) * rsqrt(var + eps) * gamma + beta, statistics in fp32 (biased variance),
// like torch.nn.LayerNorm / apex FusedLayerNorm (reference megatron/mpu/layers.py:28-36,
// eps = args.layernorm_epsilon = 1e-5, megatron/arguments.py:199).  h % 8 == 0, h <= 1024.
// mean / rstd ([rows] fp32) are optional outputs for the backward pass.
cudaError_t launch_layernorm_fwd(bool bf16, const void* x, int64_t ldx, const void* gamma,
                                 const void* beta, void* y, int64_t ldy, int rows, int h, float eps,
                                 float* mean, float* rstd, int sm_count, cudaStream_t stream);

// out[t,:] = word[ids[t],:] + pos[t % seq,:] (+ type[types[t],:]): Embedding.forward with dropout
// off (reference megatron/model/language_model.py:169-181; position ids = arange(seq),
// bert_model.py:51-58, t5_model.py:42-49).  ids / types int64 [tokens].
cudaError_t launch_embedding_fwd(bool bf16, const int64_t* ids, const int64_t* types,
                                 const void* word, const void* pos, const void* type_emb, void* out,
                                 int tokens, int seq, int h, int vocab, int num_types,
                                 cudaStream_t stream, const int32_t* pos_ids = nullptr, int max_pos = 0);

// logprob[r] = logits[r, labels[r]] - log(sum_v exp(logits[r, v])) and lse[r] = that log-sum-exp:
// the log_softmax + gather of the reader / retriever losses (reference
// tasks/openqa/e2eqa/train_e2eqa.py:82-98 and the CrossEntropyLoss at :156-160) in one pass over
// the logits, fp32 math.  labels outside [0, vocab) give logprob 0.  16-byte loads when vocab and
// ld are multiples of 8, element-wise otherwise.
cudaError_t launch_token_logprob(bool bf16, const void* logits, int64_t ld, const int64_t* labels,
                                 float* logprob, float* lse, int rows, int vocab,
                                 cudaStream_t stream);

// ---- backward (rowops_bwd.cu).  Parameter gradients accumulate into fp32 buffers (atomic adds).
// LayerNorm: dx = rstd (g - mean(g) - xhat mean(g xhat)) + dres, g = dy gamma; dgamma += sum dy xhat;
// dbeta += sum dy.  dres (optional) is the gradient arriving on the residual branch of the pre-LN
// layer (transformer.py:480-515), fused here instead of a separate add.
cudaError_t launch_layernorm_bwd(bool bf16, const void* dy, int64_t ldy, const void* x, int64_t ldx,
                                 const void* gamma, const float* mean, const float* rstd, const void* dres,
                                 int64_t ldr, void* dx, int64_t lddx, float* dgamma, float* dbeta, int rows,
                                 int h, int sm_count, cudaStream_t stream);
// out[n] += sum_m dy[m, n]  (bias gradients)
cudaError_t launch_colsum(bool bf16, const void* dy, int64_t ld, float* out, int rows, int n,
                          cudaStream_t stream);
// dlogits[r, v] = g[r] (1[v == label_r] - softmax(logits[r])_v): backward of launch_token_logprob
cudaError_t launch_token_logprob_bwd(bool bf16, const void* logits, int64_t ld, const int64_t* labels,
                                     const float* lse, const float* g, void* dlogits, int64_t ldd, int rows,
                                     int vocab, cudaStream_t stream);
// dword[ids[t]] += dx[t]; dpos[t % seq] += dx[t]; dtype_emb[types[t]] += dx[t]
cudaError_t launch_embedding_bwd(bool bf16, const void* dx, const int64_t* ids, const int64_t* types,
                                 float* dword, float* dpos, float* dtype_emb, int tokens, int seq, int h,
                                 int vocab, int num_types, cudaStream_t stream);

}  // namespace emdr2

// ---- dropout (dropout.cuh) ---------------------------------------------------------------------
#include "dropout.cuh"
namespace emdr2 {
// table[c] = B(c) for c < n (one table per seed; every dropout launch of that seed reads it)
cudaError_t launch_dropout_colhash(uint64_t seed, uint32_t* table, int n, cudaStream_t stream);
// mask[r, c] = 1 where element (r, c) is kept (test / inspection aid)
cudaError_t launch_dropout_mask(const DropoutArgs& d, uint8_t* mask, int64_t rows, int cols, cudaStream_t stream);
// out[r, :] = residual[r, :] + dropout(y[r, :])   (residual may be NULL; out may alias y): the
// bias-dropout-add of transformer.py:397-419 once the bias is in y, the embedding dropout of
// language_model.py:181, and — with y = the incoming gradient — their backward.
cudaError_t launch_dropout_add(bool bf16, const void* y, int64_t ldy, const void* residual, int64_t ldr, void* out,
                               int64_t ldo, int rows, int cols, const DropoutArgs& d, cudaStream_t stream);
}  // namespace emdr2
