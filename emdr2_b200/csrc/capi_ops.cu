// C ABI of the transformer-block operators (GEMM with fused epilogue, LayerNorm, attention, ...).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

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

// ---- optional per-kernel-kind launch timing (emdr2_ops_timing): CUDA event pairs recorded on the
// caller's stream around every launch of a kind, summed on request.  Process-wide (backward passes
// run on PyTorch's autograd worker threads), guarded by a mutex; off by default.
struct KindTimer {
  std::vector<cudaEvent_t> ev;
  size_t used = 0;
  double flops = 0.0;   // algorithmic work of the timed launches (GEMM: 2mnk; attention: 4 b h sq sk d)
};
std::atomic<bool> g_timing{false};
std::mutex g_timer_mutex;
KindTimer g_timers[EMDR2_KIND_COUNT];

struct ScopedTimer {
  KindTimer* t = nullptr;
  cudaStream_t stream;
  ScopedTimer(int kind, cudaStream_t s, double flops) : stream(s) {
    if (!g_timing.load(std::memory_order_relaxed)) return;
    g_timer_mutex.lock();
    locked = true;
    t = &g_timers[kind];
    if (t->used + 2 > t->ev.size()) {
      cudaEvent_t e0, e1;
      if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        t = nullptr;
        return;
      }
      t->ev.push_back(e0);
      t->ev.push_back(e1);
    }
    t->flops += flops;
    cudaEventRecord(t->ev[t->used], stream);
  }
  ~ScopedTimer() {
    if (t) {
      cudaEventRecord(t->ev[t->used + 1], stream);
      t->used += 2;
    }
    if (locked) g_timer_mutex.unlock();
  }
  bool locked = false;
};

}  // namespace

// 1: large forward products use the CTA-pair kernel (gemm_pair.cu); 0: always the one-CTA kernel.
static std::atomic<int> g_gemm_pair{[] {
  const char* e = getenv("EMDR2_GEMM_PAIR");
  return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
}()};

// Sixteen epilogue warps (gemm.cu, kEpi 1 / 2) for the one-CTA kernel's 16-bit outputs.  0: never; 1: where the
// epilogue has arithmetic to hide (GeLU, pre-activation output, GeLU backward) — measured +16 % on GeLU and -8 % on
// bias-only epilogues (profiles/README.md, r2B); 2: whenever eligible.
static std::atomic<int> g_gemm_wide{[] {
  const char* e = getenv("EMDR2_GEMM_WIDE");
  return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
}()};

// > 0: persistent GEMM grids use at most this many CTAs (SMs).  A trainer that overlaps NCCL all-reduces with the
// backward pass sets it to sm_count - (NCCL's CTAs): a persistent grid that asks for ALL SMs while a few are held by a
// collective kernel runs its last CTAs as a second wave and takes up to twice as long.
static std::atomic<int> g_gemm_max_ctas{0};

extern "C" {

int emdr2_ops_set_option(const char* name, int64_t value) {
  if (!name) return fail(EMDR2_EINVAL, "NULL option name");
  if (strcmp(name, "gemm_max_ctas") == 0) {
    if (value < 0 || value > 4096) return fail(EMDR2_EINVAL, "gemm_max_ctas takes 0 (all SMs) or a CTA count");
    g_gemm_max_ctas.store(static_cast<int>(value), std::memory_order_relaxed);
    return EMDR2_OK;
  }
  if (strcmp(name, "gemm_wide") == 0) {
    if (value < 0 || value > 2) return fail(EMDR2_EINVAL, "gemm_wide takes 0 (off), 1 (auto) or 2 (whenever eligible)");
    g_gemm_wide.store(static_cast<int>(value), std::memory_order_relaxed);
    return EMDR2_OK;
  }
  if (strcmp(name, "gemm_pair") == 0) {
    if (value < 0 || value > 2) return fail(EMDR2_EINVAL, "gemm_pair takes 0 (off), 1 (on) or 2 (auto)");
    g_gemm_pair.store(static_cast<int>(value), std::memory_order_relaxed);
    return EMDR2_OK;
  }
  return fail(EMDR2_EINVAL, "unknown option '%s'", name);
}

int emdr2_ops_get_option(const char* name, int64_t* out_value) {
  if (!name || !out_value) return fail(EMDR2_EINVAL, "NULL argument");
  if (strcmp(name, "gemm_wide") == 0) {
    *out_value = g_gemm_wide.load(std::memory_order_relaxed);
    return EMDR2_OK;
  }
  if (strcmp(name, "gemm_pair") == 0) {
    *out_value = g_gemm_pair.load(std::memory_order_relaxed);
    return EMDR2_OK;
  }
  if (strcmp(name, "gemm_max_ctas") == 0) {
    *out_value = g_gemm_max_ctas.load(std::memory_order_relaxed);
    return EMDR2_OK;
  }
  return fail(EMDR2_EINVAL, "unknown option '%s'", name);
}

int emdr2_gemm_ex(int dtype, const void* a, int64_t lda, int a_mn, const void* b, int64_t ldb, int b_mn,
                  void* d, int64_t ldd, const void* bias, const void* aux, int64_t ld_aux, void* preact,
                  int64_t ld_preact, int m, int n, int k, int flags, int splits, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (m < 0 || n < 0 || k < 0) return fail(EMDR2_EINVAL, "negative GEMM shape m=%d n=%d k=%d", m, n, k);
  if (m == 0 || n == 0) return EMDR2_OK;
  if (k == 0) return fail(EMDR2_EINVAL, "k=0 is not supported");
  const bool accum = (flags & EMDR2_GEMM_ACCUM_F32) != 0;
  const int out_mult = accum ? 4 : 8;   // 16-byte rows: 8 16-bit or 4 fp32 elements
  if ((n % 8) || (lda % 8) || (ldb % 8) || (ldd % out_mult))
    return fail(EMDR2_EINVAL, "n and the leading dimensions must be multiples of 8 (16-byte rows)");
  if (lda < (a_mn ? m : k) || ldb < (b_mn ? n : k) || ldd < n)
    return fail(EMDR2_EINVAL, "leading dimension smaller than the row length");
  if (!a || !b || !d) return fail(EMDR2_EINVAL, "NULL operand pointer");
  if (!aligned16(a) || !aligned16(b) || !aligned16(d)) return fail(EMDR2_EINVAL, "operands must be 16-byte aligned");
  if (a_mn && !b_mn) return fail(EMDR2_EUNSUPPORTED, "a MN-major with b K-major is not built");
  const int known = EMDR2_GEMM_BIAS | EMDR2_GEMM_GELU | EMDR2_GEMM_RESIDUAL | EMDR2_GEMM_ACCUM_F32 |
                    EMDR2_GEMM_GELU_BWD | EMDR2_GEMM_PREACT;
  if (flags & ~known) return fail(EMDR2_EINVAL, "unknown GEMM epilogue flags 0x%x", flags);
  if ((flags & EMDR2_GEMM_BIAS) && (!bias || !aligned16(bias)))
    return fail(EMDR2_EINVAL, "EMDR2_GEMM_BIAS needs a 16-byte aligned bias pointer");
  const bool has_aux = (flags & (EMDR2_GEMM_RESIDUAL | EMDR2_GEMM_GELU_BWD)) != 0;
  if ((flags & EMDR2_GEMM_RESIDUAL) && (flags & EMDR2_GEMM_GELU_BWD))
    return fail(EMDR2_EINVAL, "RESIDUAL and GELU_BWD both read the aux tile: pick one");
  if (has_aux && (!aux || !aligned16(aux) || (ld_aux % 8) || ld_aux < n))
    return fail(EMDR2_EINVAL, "RESIDUAL/GELU_BWD need a 16-byte aligned aux [m, n] with ld %% 8 == 0, ld >= n");
  if ((flags & EMDR2_GEMM_PREACT) && (!preact || !aligned16(preact) || (ld_preact % 8) || ld_preact < n))
    return fail(EMDR2_EINVAL, "PREACT needs a 16-byte aligned output [m, n] with ld %% 8 == 0, ld >= n");
  if (accum && (flags & ~(EMDR2_GEMM_ACCUM_F32)))
    return fail(EMDR2_EINVAL, "ACCUM_F32 cannot be combined with other epilogue flags");
  if (splits < 1) splits = 1;
  if (splits > 1 && !accum) return fail(EMDR2_EINVAL, "split-K needs EMDR2_GEMM_ACCUM_F32");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  static bool prepared[64] = {};
  if (!prepared[info.device]) {
    CUDA_TRY(emdr2::gemm_prepare());
    CUDA_TRY(emdr2::gemm_pair_prepare());
    prepared[info.device] = true;
  }
  CUtensorMap ta, tb, td, tr, tp;
  // Large K-major products with a 16-bit output run on CTA pairs (gemm_pair.cu): 256 x 256 tiles, at
  // least two per pair so the persistent pipeline has something to overlap.
  const int64_t pair_tiles = static_cast<int64_t>((m + 255) / 256) * ((n + emdr2::kGemmBN - 1) / emdr2::kGemmBN);
  const int pair_mode = g_gemm_pair.load(std::memory_order_relaxed);   // 0 never, 1 whenever eligible, 2 auto
  // auto: where the pair kernel is measured to win (profiles/README.md, r1g): residual / aux epilogues
  // over >= 100 k rows (+3.5 % at K = 3072, +8 % at K = 768); elsewhere the one-CTA kernel is as fast or faster
  const bool pair_wins = has_aux && m >= 100000;
  const bool use_pair = (pair_mode == 1 || (pair_mode == 2 && pair_wins)) && !a_mn && !b_mn && !accum &&
                        splits == 1 && info.sm_count >= 2 && pair_tiles >= 2 * (info.sm_count / 2);
  // write-only epilogues of the one-CTA kernel: sixteen epilogue warps, 32-column output boxes
  const int wide_mode = g_gemm_wide.load(std::memory_order_relaxed);
  const bool wide_ok = !use_pair && !a_mn && !accum && splits == 1 && !(has_aux && (flags & EMDR2_GEMM_PREACT)) &&
                       (wide_mode == 2 || (wide_mode == 1 && (flags & (EMDR2_GEMM_GELU | EMDR2_GEMM_PREACT |
                                                                       EMDR2_GEMM_GELU_BWD)) != 0));
  const int epi = !wide_ok ? 0 : has_aux ? 2 : 1;
  const bool wide = epi == 1;   // 32-column output boxes
  // K-major operand: [rows, k] with box rows x 64; MN-major: [k, rows] with 64 x 64 boxes
  rc = a_mn ? make_tmap_2d(&ta, dtype, a, k, m, lda, 64) : make_tmap_2d(&ta, dtype, a, m, k, lda, emdr2::kGemmBM);
  if (rc != EMDR2_OK) return rc;
  rc = b_mn ? make_tmap_2d(&tb, dtype, b, k, n, ldb, 64)
            : make_tmap_2d(&tb, dtype, b, n, k, ldb, use_pair ? emdr2::kGemmBN / 2 : emdr2::kGemmBN);
  if (rc != EMDR2_OK) return rc;
  if (accum) {
    td = tb;   // no 16-bit output map needed; any valid descriptor fills the unused slots
  } else if ((rc = make_tmap_2d(&td, dtype, d, m, n, ldd, emdr2::kGemmBM, wide)) != EMDR2_OK) {
    return rc;
  }
  tr = td;
  tp = td;
  if (has_aux && (rc = make_tmap_2d(&tr, dtype, aux, m, n, ld_aux, emdr2::kGemmBM)) != EMDR2_OK) return rc;
  if ((flags & EMDR2_GEMM_PREACT) &&
      (rc = make_tmap_2d(&tp, dtype, preact, m, n, ld_preact, emdr2::kGemmBM, wide)) != EMDR2_OK)
    return rc;
  emdr2::GemmArgs ga;
  ga.M = m;
  ga.N = n;
  ga.K = k;
  ga.tiles_m = (m + emdr2::kGemmBM - 1) / emdr2::kGemmBM;
  ga.tiles_n = (n + emdr2::kGemmBN - 1) / emdr2::kGemmBN;
  ga.idesc = emdr2::ptx::instr_desc_f16(dtype == EMDR2_DTYPE_BF16 ? 1 : 0, emdr2::kGemmBM, emdr2::kGemmBN,
                                        a_mn ? 1 : 0, b_mn ? 1 : 0);
  ga.flags = static_cast<uint32_t>(flags);
  const uint32_t num_kb = (k + emdr2::kGemmBK - 1) / emdr2::kGemmBK;
  if (static_cast<uint32_t>(splits) > num_kb) splits = static_cast<int>(num_kb);
  ga.kb_per_split = (num_kb + splits - 1) / splits;
  ga.splits = (num_kb + ga.kb_per_split - 1) / ga.kb_per_split;
  ga.ldd32 = static_cast<uint32_t>(ldd);
  ga.bias = bias;
  ga.out32 = accum ? static_cast<float*>(d) : nullptr;
  const uint32_t work = ga.tiles_m * ga.tiles_n * ga.splits;
  const int cap = g_gemm_max_ctas.load(std::memory_order_relaxed);
  const int sms = (cap > 0 && cap < info.sm_count) ? cap : info.sm_count;
  const int grid = static_cast<int>(work < static_cast<uint32_t>(sms) ? work : sms);
  ScopedTimer timer(EMDR2_KIND_GEMM, static_cast<cudaStream_t>(cuda_stream), 2.0 * m * n * k);
  if (use_pair) {
    ga.idesc = emdr2::ptx::instr_desc_f16(dtype == EMDR2_DTYPE_BF16 ? 1 : 0, 2 * emdr2::kGemmBM, emdr2::kGemmBN);
    const int pairs = sms / 2;
    CUDA_TRY(emdr2::launch_gemm_pair(ta, tb, td, tr, tp, ga, dtype == EMDR2_DTYPE_BF16, 2 * pairs,
                                     static_cast<cudaStream_t>(cuda_stream)));
    return EMDR2_OK;
  }
  CUDA_TRY(emdr2::launch_gemm(ta, tb, td, tr, tp, ga, dtype == EMDR2_DTYPE_BF16, a_mn != 0, b_mn != 0, epi, grid,
                              static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_gemm(int dtype, const void* a, int64_t lda, const void* b, int64_t ldb, void* d,
               int64_t ldd, const void* bias, const void* residual, int64_t ldr, int m, int n,
               int k, int flags, void* cuda_stream) {
  if (flags & ~(EMDR2_GEMM_BIAS | EMDR2_GEMM_GELU | EMDR2_GEMM_RESIDUAL))
    return fail(EMDR2_EINVAL, "unknown GEMM epilogue flags 0x%x", flags);
  return emdr2_gemm_ex(dtype, a, lda, 0, b, ldb, 0, d, ldd, bias, residual, ldr, nullptr, 0, m, n, k, flags, 1,
                       cuda_stream);
}

static int check_dropout(float p, const uint32_t* colhash) {
  if (!(p >= 0.f) || p >= 1.f) return fail(EMDR2_EINVAL, "dropout probability %g is not in [0, 1)", p);
  if (p > 0.f && (!colhash || !aligned16(colhash)))
    return fail(EMDR2_EINVAL, "dropout needs a 16-byte aligned column-hash table (emdr2_dropout_colhash)");
  return EMDR2_OK;
}

int emdr2_attention_fwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                        const void* v, int64_t ldv, void* o, int64_t ldo, int batch, int heads,
                        int sq, int sk, const uint8_t* q_pad, const uint8_t* k_pad,
                        const uint8_t* q_live, const uint8_t* k_live, int causal, float scale,
                        float* lse, void* cuda_stream) {
  return emdr2_attention_fwd_dropout(dtype, q, ldq, k, ldk, v, ldv, o, ldo, batch, heads, sq, sk, q_pad, k_pad,
                                     q_live, k_live, causal, scale, lse, 0.f, 0, 0, nullptr, cuda_stream);
}

int emdr2_attention_fwd_dropout(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                const void* v, int64_t ldv, void* o, int64_t ldo, int batch, int heads,
                                int sq, int sk, const uint8_t* q_pad, const uint8_t* k_pad,
                                const uint8_t* q_live, const uint8_t* k_live, int causal, float scale,
                                float* lse, float p, uint64_t seed, uint64_t offset, const uint32_t* colhash,
                                void* cuda_stream) {
  if (check_dropout(p, colhash) != EMDR2_OK) return EMDR2_EINVAL;
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
    CUDA_TRY(emdr2::attention_persistent_prepare());
    prepared[info.device] = true;
  }
  static const bool legacy = [] {
    const char* e = getenv("EMDR2_ATTN_IMPL");
    return e && strcmp(e, "legacy") == 0;
  }();
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
  aa.q_live = q_live;
  aa.k_live = k_live;
  aa.lse = lse;
  aa.drop = emdr2::make_dropout_args(p, seed, offset, colhash);
  if (aa.drop.threshold && legacy) return fail(EMDR2_EUNSUPPORTED, "the legacy attention kernel has no dropout");
  ScopedTimer timer(EMDR2_KIND_ATTENTION, static_cast<cudaStream_t>(cuda_stream),
                    4.0 * batch * heads * sq * static_cast<double>(sk) * emdr2::kAttnHeadDim);
  if (legacy)
    emdr2::launch_attention_fwd(tq, tk, tv, to, aa, fmt == 1, static_cast<cudaStream_t>(cuda_stream));
  else
    emdr2::launch_attention_fwd_persistent(tq, tk, tv, to, aa, fmt == 1, info.sm_count,
                                           static_cast<cudaStream_t>(cuda_stream));
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
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_layernorm_fwd(dtype == EMDR2_DTYPE_BF16, x, ldx, gamma, beta, y, ldy, rows, h,
                                       eps, mean, rstd, info.sm_count, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_embedding_fwd(int dtype, const int64_t* ids, const int64_t* types, const void* word,
                        const void* pos, const void* type_emb, void* out, int tokens, int seq, int h,
                        int vocab, int num_types, void* cuda_stream) {
  return emdr2_embedding_fwd_pos(dtype, ids, types, word, pos, type_emb, out, tokens, seq, h, vocab, num_types,
                                 nullptr, seq, cuda_stream);
}

int emdr2_embedding_fwd_pos(int dtype, const int64_t* ids, const int64_t* types, const void* word,
                            const void* pos, const void* type_emb, void* out, int tokens, int seq, int h,
                            int vocab, int num_types, const int32_t* pos_ids, int max_pos, void* cuda_stream) {
  if (pos_ids && max_pos < 1) return fail(EMDR2_EINVAL, "explicit positions need the position table's row count");
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
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_embedding_fwd(dtype == EMDR2_DTYPE_BF16, ids, types, word, pos, type_emb, out,
                                       tokens, seq, h, vocab, num_types,
                                       static_cast<cudaStream_t>(cuda_stream), pos_ids, max_pos));
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
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_token_logprob(dtype == EMDR2_DTYPE_BF16, logits, ld, labels, logprob, lse, rows,
                                       vocab, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_ops_timing(int enable) {
  std::lock_guard<std::mutex> guard(g_timer_mutex);
  g_timing = enable != 0;
  for (int k = 0; k < EMDR2_KIND_COUNT; ++k) {
    g_timers[k].used = 0;
    g_timers[k].flops = 0.0;
  }
  return EMDR2_OK;
}

int emdr2_ops_timing_add_flops(int kind, double flops) {
  if (kind < 0 || kind >= EMDR2_KIND_COUNT) return fail(EMDR2_EINVAL, "unknown kernel kind %d", kind);
  if (!g_timing.load(std::memory_order_relaxed)) return EMDR2_OK;
  std::lock_guard<std::mutex> guard(g_timer_mutex);
  g_timers[kind].flops += flops;
  return EMDR2_OK;
}

int emdr2_ops_timing_read(int kind, int64_t* out_ns, int64_t* out_launches, double* out_flops) {
  if (kind < 0 || kind >= EMDR2_KIND_COUNT) return fail(EMDR2_EINVAL, "unknown kernel kind %d", kind);
  std::lock_guard<std::mutex> guard(g_timer_mutex);
  KindTimer& t = g_timers[kind];
  double total_ms = 0.0;
  for (size_t i = 0; i + 1 < t.used; i += 2) {
    float ms = 0.f;
    CUDA_TRY(cudaEventSynchronize(t.ev[i + 1]));
    CUDA_TRY(cudaEventElapsedTime(&ms, t.ev[i], t.ev[i + 1]));
    total_ms += ms;
  }
  if (out_ns) *out_ns = static_cast<int64_t>(total_ms * 1e6);
  if (out_launches) *out_launches = static_cast<int64_t>(t.used / 2);
  if (out_flops) *out_flops = t.flops;
  t.used = 0;
  t.flops = 0.0;
  return EMDR2_OK;
}

int emdr2_attention_bwd(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                        int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo, void* dq,
                        int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int batch, int heads,
                        int sq, int sk, const uint8_t* q_pad, const uint8_t* k_pad, const uint8_t* q_live,
                        const uint8_t* k_live, int causal, float scale, const float* lse, float* dvec_ws,
                        void* cuda_stream) {
  return emdr2_attention_bwd_dropout(dtype, q, ldq, k, ldk, v, ldv, o, ldo, dout, lddo, dq, lddq, dk, lddk, dv, lddv,
                                     batch, heads, sq, sk, q_pad, k_pad, q_live, k_live, causal, scale, lse, dvec_ws,
                                     0.f, 0, 0, nullptr, cuda_stream);
}

int emdr2_attention_bwd_dropout(int dtype, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                int64_t ldv, const void* o, int64_t ldo, const void* dout, int64_t lddo, void* dq,
                                int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int batch, int heads,
                                int sq, int sk, const uint8_t* q_pad, const uint8_t* k_pad, const uint8_t* q_live,
                                const uint8_t* k_live, int causal, float scale, const float* lse, float* dvec_ws,
                                float p, uint64_t seed, uint64_t offset, const uint32_t* colhash,
                                void* cuda_stream) {
  if (check_dropout(p, colhash) != EMDR2_OK) return EMDR2_EINVAL;
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (batch < 0 || heads < 1 || sq < 0 || sk < 1)
    return fail(EMDR2_EINVAL, "bad attention shape batch=%d heads=%d sq=%d sk=%d", batch, heads, sq, sk);
  if (batch == 0 || sq == 0) return EMDR2_OK;
  if (heads > 65535 || batch > 65535) return fail(EMDR2_EINVAL, "batch and heads must be <= 65535");
  const void* ptrs[] = {q, k, v, o, dout, dq, dk, dv};
  for (const void* p : ptrs)
    if (!p || !aligned16(p)) return fail(EMDR2_EINVAL, "NULL or misaligned tensor passed to emdr2_attention_bwd");
  if (!lse || !dvec_ws) return fail(EMDR2_EINVAL, "lse and the dvec workspace ([batch, heads, sq] fp32) are required");
  const int64_t width = static_cast<int64_t>(heads) * emdr2::kAttnHeadDim;
  const int64_t lds[] = {ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv};
  for (int64_t ld : lds)
    if ((ld % 8) || ld < width) return fail(EMDR2_EINVAL, "row pitches must be multiples of 8 and >= heads*64");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  static bool prepared[64] = {};
  if (!prepared[info.device]) {
    CUDA_TRY(emdr2::attention_bwd_prepare());
    prepared[info.device] = true;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const bool bf16 = dtype == EMDR2_DTYPE_BF16;
  ScopedTimer timer(EMDR2_KIND_ATTENTION, stream,
                    14.0 * batch * heads * sq * static_cast<double>(sk) * emdr2::kAttnHeadDim);
  CUDA_TRY(emdr2::launch_attention_bwd_prep(bf16, dout, lddo, o, ldo, dvec_ws, batch, heads, sq, stream));
  emdr2::AttnBwdMaps mp;
  struct { CUtensorMap* m; const void* p; int rows; int64_t ld; uint32_t box; } specs[] = {
      {&mp.q128, q, sq, ldq, 128}, {&mp.do128, dout, sq, lddo, 128}, {&mp.dq128, dq, sq, lddq, 128},
      {&mp.k64, k, sk, ldk, 64},   {&mp.v64, v, sk, ldv, 64},        {&mp.k128, k, sk, ldk, 128},
      {&mp.v128, v, sk, ldv, 128}, {&mp.dk128, dk, sk, lddk, 128},   {&mp.dv128, dv, sk, lddv, 128},
      {&mp.q64, q, sq, ldq, 64},   {&mp.do64, dout, sq, lddo, 64}};
  for (auto& sp : specs)
    if ((rc = make_tmap_3d(sp.m, dtype, sp.p, batch, sp.rows, width, sp.ld, sp.box)) != EMDR2_OK) return rc;
  emdr2::AttnBwdArgs aa;
  aa.batch = batch;
  aa.heads = heads;
  aa.sq = sq;
  aa.sk = sk;
  aa.causal = causal ? 1u : 0u;
  const int fmt = bf16 ? 1 : 0;
  aa.idesc_s = emdr2::ptx::instr_desc_f16(fmt, 128, 64);
  aa.idesc_o = emdr2::ptx::instr_desc_f16(fmt, 128, emdr2::kAttnHeadDim, 0, 1);
  aa.scale = scale;
  aa.scale_log2 = scale * 1.4426950408889634f;
  aa.q_pad = q_pad;
  aa.k_pad = k_pad;
  aa.q_live = q_live;
  aa.k_live = k_live;
  aa.lse = lse;
  aa.dvec = dvec_ws;
  aa.drop = emdr2::make_dropout_args(p, seed, offset, colhash);
  CUDA_TRY(emdr2::launch_attention_bwd(mp, aa, bf16, stream));
  return EMDR2_OK;
}

int emdr2_layernorm_bwd(int dtype, const void* dy, int64_t ldy, const void* x, int64_t ldx, const void* gamma,
                        const float* mean, const float* rstd, const void* dres, int64_t ldr, void* dx,
                        int64_t lddx, float* dgamma, float* dbeta, int rows, int h, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (rows < 0 || h < 8 || (h % 8) || h > 1024)
    return fail(EMDR2_EINVAL, "layernorm_bwd needs rows >= 0 and h a multiple of 8 in [8, 1024]");
  if (rows == 0) return EMDR2_OK;
  if (!dy || !x || !gamma || !mean || !rstd || !dx) return fail(EMDR2_EINVAL, "NULL pointer passed to emdr2_layernorm_bwd");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(gamma) || !aligned16(dx) || (dres && !aligned16(dres)) ||
      (ldy % 8) || (ldx % 8) || (lddx % 8) || (dres && (ldr % 8)))
    return fail(EMDR2_EINVAL, "layernorm_bwd operands must be 16-byte aligned with row pitches %% 8 == 0");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_layernorm_bwd(dtype == EMDR2_DTYPE_BF16, dy, ldy, x, ldx, gamma, mean, rstd, dres, ldr, dx,
                                       lddx, dgamma, dbeta, rows, h, info.sm_count,
                                       static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_colsum(int dtype, const void* dy, int64_t ld, float* out, int rows, int n, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (rows < 0 || n < 0 || (n % 8) || (ld % 8) || ld < n) return fail(EMDR2_EINVAL, "colsum needs n %% 8 == 0, ld %% 8 == 0, ld >= n");
  if (rows == 0 || n == 0) return EMDR2_OK;
  if (!dy || !out || !aligned16(dy)) return fail(EMDR2_EINVAL, "NULL or misaligned pointer passed to emdr2_colsum");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_colsum(dtype == EMDR2_DTYPE_BF16, dy, ld, out, rows, n, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_token_logprob_bwd(int dtype, const void* logits, int64_t ld, const int64_t* labels, const float* lse,
                            const float* g, void* dlogits, int64_t ldd, int rows, int vocab, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (rows < 0 || vocab < 8 || (vocab % 8) || (ld % 8) || (ldd % 8) || ld < vocab || ldd < vocab)
    return fail(EMDR2_EINVAL, "token_logprob_bwd needs vocab %% 8 == 0 and row pitches %% 8 == 0, >= vocab");
  if (rows == 0) return EMDR2_OK;
  if (!logits || !labels || !lse || !g || !dlogits || !aligned16(logits) || !aligned16(dlogits))
    return fail(EMDR2_EINVAL, "NULL or misaligned pointer passed to emdr2_token_logprob_bwd");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_token_logprob_bwd(dtype == EMDR2_DTYPE_BF16, logits, ld, labels, lse, g, dlogits, ldd, rows,
                                           vocab, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_embedding_bwd(int dtype, const void* dx, const int64_t* ids, const int64_t* types, float* dword,
                        float* dpos, float* dtype_emb, int tokens, int seq, int h, int vocab, int num_types,
                        void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (tokens < 0 || seq < 1 || h < 8 || (h % 8) || vocab < 1)
    return fail(EMDR2_EINVAL, "bad embedding shape tokens=%d seq=%d h=%d vocab=%d", tokens, seq, h, vocab);
  if (tokens == 0) return EMDR2_OK;
  if (!dx || !ids || !aligned16(dx)) return fail(EMDR2_EINVAL, "NULL or misaligned pointer passed to emdr2_embedding_bwd");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_embedding_bwd(dtype == EMDR2_DTYPE_BF16, dx, ids, types, dword, dpos, dtype_emb, tokens, seq,
                                       h, vocab, num_types, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_dropout_colhash(uint64_t seed, uint32_t* dev_table, int n, void* cuda_stream) {
  if (n < 0 || (n > 0 && !dev_table)) return fail(EMDR2_EINVAL, "bad column-hash table (n=%d)", n);
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  CUDA_TRY(emdr2::launch_dropout_colhash(seed, dev_table, n, static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_dropout_mask(float p, uint64_t seed, uint64_t offset, const uint32_t* colhash, uint8_t* dev_mask,
                       int64_t rows, int cols, void* cuda_stream) {
  if (!(p > 0.f) || p >= 1.f || !colhash || !dev_mask || rows < 0 || cols < 0)
    return fail(EMDR2_EINVAL, "emdr2_dropout_mask needs 0 < p < 1, a column-hash table and an output buffer");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  CUDA_TRY(emdr2::launch_dropout_mask(emdr2::make_dropout_args(p, seed, offset, colhash), dev_mask, rows, cols,
                                      static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_dropout_add(int dtype, const void* y, int64_t ldy, const void* residual, int64_t ldr, void* out,
                      int64_t ldo, int rows, int cols, float p, uint64_t seed, uint64_t offset,
                      const uint32_t* colhash, void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (!(p > 0.f) || p >= 1.f) return fail(EMDR2_EINVAL, "emdr2_dropout_add needs 0 < p < 1 (got %g)", p);
  if (check_dropout(p, colhash) != EMDR2_OK) return EMDR2_EINVAL;
  if (rows < 0 || cols < 0 || (cols % 8) || (ldy % 8) || (ldo % 8) || ldy < cols || ldo < cols ||
      (residual && ((ldr % 8) || ldr < cols)))
    return fail(EMDR2_EINVAL, "dropout_add needs cols %% 8 == 0 and row pitches %% 8 == 0, >= cols");
  if (rows == 0 || cols == 0) return EMDR2_OK;
  if (!y || !out || !aligned16(y) || !aligned16(out) || (residual && !aligned16(residual)))
    return fail(EMDR2_EINVAL, "NULL or misaligned pointer passed to emdr2_dropout_add");
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  ScopedTimer timer(EMDR2_KIND_ROWOP, static_cast<cudaStream_t>(cuda_stream), 0.0);
  CUDA_TRY(emdr2::launch_dropout_add(dtype == EMDR2_DTYPE_BF16, y, ldy, residual, ldr, out, ldo, rows, cols,
                                     emdr2::make_dropout_args(p, seed, offset, colhash),
                                     static_cast<cudaStream_t>(cuda_stream)));
  return EMDR2_OK;
}

int emdr2_attention_varlen_fwd(int dtype, const void* q, int64_t ldq, int64_t q_rows, const void* k, int64_t ldk,
                               const void* v, int64_t ldv, int64_t k_rows, void* o, int64_t ldo, int64_t o_rows,
                               int heads, const int32_t* dev_items, int n_items, float scale, float* lse,
                               void* cuda_stream) {
  if (dtype != EMDR2_DTYPE_FP16 && dtype != EMDR2_DTYPE_BF16)
    return fail(EMDR2_EINVAL, "dtype %d is not EMDR2_DTYPE_FP16/BF16", dtype);
  if (heads < 1 || n_items < 0 || q_rows < 0 || k_rows < 0 || o_rows < 0)
    return fail(EMDR2_EINVAL, "bad varlen attention shape heads=%d items=%d", heads, n_items);
  if (n_items == 0) return EMDR2_OK;
  if (q_rows == 0 || k_rows == 0 || o_rows == 0) return fail(EMDR2_EINVAL, "work items over empty matrices");
  if (q_rows > 0x7fffffff || k_rows > 0x7fffffff || o_rows > 0x7fffffff)
    return fail(EMDR2_EINVAL, "packed matrices are limited to 2^31 - 1 rows");
  if (!q || !k || !v || !o || !dev_items) return fail(EMDR2_EINVAL, "NULL pointer passed to emdr2_attention_varlen_fwd");
  if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o) || !aligned16(dev_items))
    return fail(EMDR2_EINVAL, "q/k/v/o and the item list must be 16-byte aligned");
  const int64_t width = static_cast<int64_t>(heads) * emdr2::kAttnHeadDim;
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8) || ldq < width || ldk < width || ldv < width || ldo < width)
    return fail(EMDR2_EINVAL, "row pitches must be multiples of 8 and >= heads*64 = %lld", static_cast<long long>(width));
  DeviceInfo info;
  int rc = require_b200(&info);
  if (rc != EMDR2_OK) return rc;
  static bool prepared[64] = {};
  if (!prepared[info.device]) {
    CUDA_TRY(emdr2::attention_varlen_prepare());
    prepared[info.device] = true;
  }
  CUtensorMap tq, tk, tv, to;
  if ((rc = make_tmap_2d(&tq, dtype, q, q_rows, width, ldq, emdr2::kAttnBQ)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_2d(&tk, dtype, k, k_rows, width, ldk, emdr2::kAttnBK)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_2d(&tv, dtype, v, k_rows, width, ldv, emdr2::kAttnBK)) != EMDR2_OK) return rc;
  if ((rc = make_tmap_2d(&to, dtype, o, o_rows, width, ldo, emdr2::kAttnBQ)) != EMDR2_OK) return rc;
  emdr2::AttnVarlenArgs aa;
  aa.items = reinterpret_cast<const emdr2::AttnVarlenItem*>(dev_items);
  aa.n_items = static_cast<uint32_t>(n_items);
  const int fmt = dtype == EMDR2_DTYPE_BF16 ? 1 : 0;
  aa.idesc_s = emdr2::ptx::instr_desc_f16(fmt, emdr2::kAttnBQ, emdr2::kAttnBK);
  aa.idesc_o = emdr2::ptx::instr_desc_f16(fmt, emdr2::kAttnBQ, emdr2::kAttnHeadDim, 0, 1);
  aa.scale_log2 = scale * 1.4426950408889634f;
  aa.out = o;
  aa.ldo = ldo;
  aa.lse = lse;
  static const uint32_t sched = [] {
    const char* e = getenv("EMDR2_VARLEN_SCHED");
    return (e && e[0] == '1') ? 1u : 0u;
  }();
  aa.sched = sched;
  // algorithmic work is not known here without reading the items back: callers that time this kind pass
  // through emdr2_ops_timing and account flops themselves; count the launch with zero flops
  ScopedTimer timer(EMDR2_KIND_ATTENTION, static_cast<cudaStream_t>(cuda_stream), 0.0);
  emdr2::launch_attention_varlen(tq, tk, tv, to, aa, fmt == 1, info.sm_count, static_cast<cudaStream_t>(cuda_stream));
  CUDA_TRY(cudaGetLastError());
  return EMDR2_OK;
}

}  // extern "C"
