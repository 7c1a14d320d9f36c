// C ABI of the transformer-block operators (GEMM with fused epilogue, LayerNorm, attention, ...).
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/emdr2_b200.h"
#include "capi_common.cuh"
#include "attention.cuh"
#include "gemm.cuh"
#include "rowops.cuh"
#include "ptx.cuh"

namespace {

using emdr2::capi::DeviceInfo;
using emdr2::capi::fail;
using emdr2::capi::make_tmap_2d;
using emdr2::capi::make_tmap_3d;

int require_b200(DeviceInfo* info) {
  int rc = emdr2::capi::current_device_info(info);
  if (rc != EMDR2_OK) return rc;
  if (info->major != 10)
    return fail(EMDR2_EUNSUPPORTED, "device %d is sm_%d%d; this library contains sm_100a (B200) code only",
                info->device, info->major, info->minor);
  return EMDR2_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int emdr2_gemm(int dtype, const void* a, int64_t lda, const void* b, int64_t ldb, void* d,
               int64_t ldd, const void* bias, const void* residual, int64_t ldr, int m, int n,
               int k, int flags, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (m < 0 || n < 0 || k < 0) return fail(EMDR2_EINVAL, "negative GEMM shape m=%d n=%d k=%d", m, n, k);
  if (m == 0 || n == 0) return EMDR2_OK;
  if (k == 0) return fail(EMDR2_EINVAL, "k=0 is not supported");
  if ((n % 8) || (k % 8) || (lda % 8) || (ldb % 8) || (ldd % 8))
    return fail(EMDR2_EINVAL, "n, k and the leading dimensions must be multiples of 8 (16-byte rows)");
  if (lda < k || ldb < k || ldd < n) return fail(EMDR2_EINVAL, "leading dimension smaller than the row length");
  if (!a || !b || !d) return fail(EMDR2_EINVAL, "NULL operand pointer");
  if (!aligned16(a) || !aligned16(b) || !aligned16(d)) return fail(EMDR2_EINVAL, "operands must be 16-byte aligned");
  if (flags & ~(EMDR2_GEMM_BIAS | EMDR2_GEMM_GELU | EMDR2_GEMM_RESIDUAL))
    return fail(EMDR2_EINVAL, "unknown GEMM epilogue flags 0x%x", flags);
  if ((flags & EMDR2_GEMM_BIAS) && (!bias || !aligned16(bias)))
    return fail(EMDR2_EINVAL, "EMDR2_GEMM_BIAS needs a 16-byte aligned bias pointer");
  if ((flags & EMDR2_GEMM_RESIDUAL) && (!residual || !aligned16(residual) || (ldr % 8) || ldr < n))
    return fail(EMDR2_EINVAL, "EMDR2_GEMM_RESIDUAL needs a 16-byte aligned residual with ldr %% 8 == 0, ldr >= n");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  static bool prepared[64] = {};
  if (!prepared[info.device]) {
    CUDA_TRY(emdr2::gemm_prepare());
    prepared[info.device] = true;
  }
  CUtensorMap ta, tb, td, tr;
  if ((rc = make_tmap_2d(&ta, dtype, a, m, k, lda, emdr2::kGemmBM)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_2d(&tb, dtype, b, n, k, ldb, emdr2::kGemmBN)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_2d(&td, dtype, d, m, n, ldd, emdr2::kGemmBM)) != EMDR2_OK) return rc;
  tr = td;
  if ((flags & EMDR2_GEMM_RESIDUAL) &&
      (rc = make_tmap_2d(&tr, dtype, residual, m, n, ldr, emdr2::kGemmBM)) != EMDR2_OK)
    return rc;
  emdr2::GemmArgs ga;
  ga.M = m;
  ga.N = n;
  ga.K = k;
  ga.tiles_m = (m + emdr2::kGemmBM - 1) / emdr2::kGemmBM;
  ga.tiles_n = (n + emdr2::kGemmBN - 1) / emdr2::kGemmBN;
  ga.idesc = emdr2::ptx::instr_desc_f16(dtype == EMDR2_DTYPE_BF16 ? 1 : 0, emdr2::kGemmBM, emdr2::kGemmBN);
  ga.flags = static_cast<uint32_t>(flags);
  ga.bias = bias;
  const uint32_t tiles = ga.tiles_m * ga.tiles_n;
  const int grid = static_cast<int>(tiles < static_cast<uint32_t>(info.sm_count) ? tiles : info.sm_count);
  emdr2::launch_gemm(ta, tb, td, tr, ga, dtype == EMDR2_DTYPE_BF16, grid, static_cast<cudaStream_t>(cuda_stream));
  CUDA_TRY(cudaGetLastError());
  return EMDR2_OK;
}

int emdr2_attention_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                        const void* v, int64_t ldv, void* o, int64_t ldo, int batch, int heads,
                        int sq, int sk, const uint8_t* q_pad, const uint8_t* k_pad, int causal,
                        float scale, float* lse, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (batch < 0 || heads < 1 || sq < 0 || sk < 0)
    return fail(EMDR2_EINVAL, "bad attention shape batch=%d heads=%d sq=%d sk=%d", batch, heads, sq, sk);
  if (batch == 0 || sq == 0) return EMDR2_OK;
  if (sk == 0) return fail(EMDR2_EINVAL, "sk=0: attention over an empty key set is undefined");
  if (heads > 65535 || batch > 65535) return fail(EMDR2_EINVAL, "batch and heads must be <= 65535");
  if (!q || !k || !v || !o) return fail(EMDR2_EINVAL, "NULL q/k/v/o pointer");
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o))
    return fail(EMDR2_EINVAL, "q/k/v/o must be 16-byte aligned");
  const int64_t width = static_cast<int64_t>(heads) * emdr2::kAttnHeadDim;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || ldq < width || ldk < width || ldv < width ||
      ldo < width)
    return fail(EMDR2_EINVAL, "row pitches must be multiples of 8 and >= heads*64 = %lld",
                static_cast<long long>(width));
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  static bool prepared[64] = {};
  if (!prepared[info.device]) {
    CUDA_TRY(emdr2::attention_prepare());
    prepared[info.device] = true;
  }
  CUtensorMap tq, tk, tv, to;
  if ((rc = make_tmap_3d(&tq, dtype, q, batch, sq, width, ldq, emdr2::kAttnBQ)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_3d(&tk, dtype, k, batch, sk, width, ldk, emdr2::kAttnBK)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_3d(&tv, dtype, v, batch, sk, width, ldv, emdr2::kAttnBK)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_3d(&to, dtype, o, batch, sq, width, ldo, emdr2::kAttnBQ)) != EMDR2_OK) return rc;
  emdr2::AttnArgs aa;
  aa.batch = batch;
  aa.heads = heads;
  aa.sq = sq;
  aa.sk = sk;
  aa.causal = causal ? 1u : 0u;
  const int fmt = dtype == EMDR2_DTYPE_BF16 ? 1 : 0;
  aa.idesc_s = emdr2::ptx::instr_desc_f16(fmt, emdr2::kAttnBQ, emdr2::kAttnBK);
  aa.idesc_o = emdr2::ptx::instr_desc_f16(fmt, emdr2::kAttnBQ, emdr2::kAttnHeadDim, 0, 1);
  aa.scale_log2 = scale * 1.4426950408889634f;
  aa.q_pad = q_pad;
  aa.k_pad = k_pad;
  aa.lse = lse;
  emdr2::launch_attention_fwd(tq, tk, tv, to, aa, fmt == 1, static_cast<cudaStream_t>(cuda_stream));
  CUDA_TRY(cudaGetLastError());
  return EMDR2_OK;
}

int emdr2_layernorm_fwd(int dtype, const void* x, int64_t ldx, const void* gamma, const void* beta,
                        void* y, int64_t ldy, int rows, int h, float eps, float* mean, float* rstd,
                        void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (rows < 0 || h < 8 || (h % 8) || h > 1024)
    return fail(EMDR2_EINVAL, "layernorm needs rows >= 0 and h a multiple of 8 in [8, 1024] (rows=%d h=%d)", rows, h);
  if (rows == 0) return EMDR2_OK;
  if (!x || !gamma || !beta || !y) return fail(EMDR2_EINVAL, "NULL pointer passed to emdr2_layernorm_fwd");
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || (ldx % 8) || (ldy % 8) ||
      ldx < h || ldy < h)
    return fail(EMDR2_EINVAL, "layernorm operands must be 16-byte aligned with row pitches %% 8 == 0, >= h");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  CUDA_TRY(emdr2::launch_layernorm_fwd(dtype == EMDR2_DTYPE_BF16, x, ldx, gamma, beta, y, ldy, rows, h,
                                       eps, mean, rstd, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_embedding_fwd(int dtype, const int64_t* ids, const int64_t* types, const void* word,
                        const void* pos, const void* type_emb, void* out, int tokens, int seq, int h,
                        int vocab, int num_types, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (tokens < 0 || seq < 1 || h < 8 || (h % 8) || vocab < 1)
    return fail(EMDR2_EINVAL, "bad embedding shape tokens=%d seq=%d h=%d vocab=%d", tokens, seq, h, vocab);
  if (tokens == 0) return EMDR2_OK;
  if (!ids || !word || !pos || !out) return fail(EMDR2_EINVAL, "NULL pointer passed to emdr2_embedding_fwd");
  if (types && (!type_emb || num_types < 1))
    return fail(EMDR2_EINVAL, "token types given without a token-type table");
  if (!aligned16(word) || !aligned16(pos) || !aligned16(out) || (type_emb && !aligned16(type_emb)))
    return fail(EMDR2_EINVAL, "embedding tables and output must be 16-byte aligned");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  CUDA_TRY(emdr2::launch_embedding_fwd(dtype == EMDR2_DTYPE_BF16, ids, types, word, pos, type_emb, out,
                                       tokens, seq, h, vocab, num_types,
                                       static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_token_logprob(int dtype, const void* logits, int64_t ld, const int64_t* labels,
                        float* logprob, float* lse, int rows, int vocab, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (rows < 0 || vocab < 1 || ld < vocab)
    return fail(EMDR2_EINVAL, "token_logprob needs rows >= 0, vocab >= 1, ld >= vocab");
  if (rows == 0) return EMDR2_OK;
  if (!logits || !labels || !logprob) return fail(EMDR2_EINVAL, "NULL pointer passed to emdr2_token_logprob");
  if ((vocab % 8 == 0) && (ld % 8 == 0) && !aligned16(logits))
    return fail(EMDR2_EINVAL, "logits must be 16-byte aligned");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  CUDA_TRY(emdr2::launch_token_logprob(dtype == EMDR2_DTYPE_BF16, logits, ld, labels, logprob, lse, rows,
                                       vocab, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

}  // extern "C"
