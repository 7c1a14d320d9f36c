// Fused attention forward for the BERT towers and the T5 reader, sm_100a only, head dim 64.
//
//   O[b, i, h, :] = softmax_j( mask( scale * Q[b,i,h,:]·K[b,j,h,:] ) ) · V[b, j, h, :]
//
// Replaces ParallelAttention.forward's core (reference megatron/model/transformer.py:301-383:
// baddbmm :309 -> FusedScaleMaskSoftmax :340 (megatron/model/fused_softmax.py:116-125, unfused
// branch: masked_fill(mask, -10000) then softmax) -> dropout (p = 0 here) -> bmm :371 -> the
// [b,np,sq,hn] -> [sq,b,hp] permute :377-383).  The [b,np,sq,sk] score / probability tensors of the
// reference never exist: scores live in TMEM, probabilities in shared memory.
//
// Mask semantics are the reference's, not -inf: masked(i,j) = q_pad[b,i] | k_pad[b,j] |
// (causal & j > i) and a masked score is REPLACED by -10000.0 (bert_model.py:31-33,
// t5_model.py:28-30), so a fully masked query row attends uniformly to all sk keys.  The masks the
// reference materialises as [b, sq, sk] bool tensors (make_attention_mask_3d(x, y) < 0.5,
// megatron/data/mask_creation_utils.py:17-26; history mask :37-42) are exactly this outer-product /
// triangular form, so two byte vectors and a flag carry them.
//
// One CTA per (128-query block, head, batch).  Per 128-key block: S = Q·Kᵀ (tcgen05, fp32 in TMEM,
// double-buffered) -> 4 softmax warps, one thread per query row (online softmax in the log2
// domain) -> P (16-bit) into shared memory in the K-major SW128 operand layout -> O_j = P·V_j
// (V consumed MN-major straight from its TMA tile) -> accumulated in registers with the usual
// running-max rescale.  K/V stream through a 2-stage TMA ring.  Two CTAs are resident per SM
// (setmaxnreg moves registers from the TMA/MMA warps to the softmax warps to make that fit).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dropout.cuh"

namespace emdr2 {

constexpr int kAttnHeadDim = 64;
constexpr int kAttnBQ = 128;      // query rows per CTA
constexpr int kAttnBK = 128;      // keys per block
constexpr int kAttnStages = 2;
constexpr int kAttnThreads = 256; // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 softmax
constexpr int kAttnTileBytes = kAttnBQ * kAttnHeadDim * 2;              // 16 KiB
// 112 KiB of tiles + 256 B of barriers: two CTAs fit one SM (2 x 113 KiB), so one CTA's TMA / MMA /
// softmax latencies are hidden behind the other's; TMEM is split 256 + 256 columns.
constexpr int kAttnBarBytes = 256;
constexpr int kAttnSmemBytes = kAttnTileBytes                            // Q (reused for O)
                               + kAttnStages * 2 * kAttnTileBytes        // K/V ring
                               + 2 * kAttnTileBytes                      // P (two 64-key K blocks)
                               + kAttnBarBytes;
constexpr int kAttnTmemCols = 256;                                       // S [0,128) + O [128,192)

struct AttnArgs {
  uint32_t batch, heads, sq, sk;
  uint32_t causal;
  uint32_t idesc_s;        // M=128, N=128, A/B K-major
  uint32_t idesc_o;        // M=128, N=64,  B MN-major
  float scale_log2;        // scale * log2(e)
  const uint8_t* q_pad;    // [batch, sq] or nullptr (1 = padding token)
  const uint8_t* k_pad;    // [batch, sk] or nullptr
  const uint8_t* q_live;   // [batch, ceil(sq/128)] or nullptr: 0 = query block is all padding (store zeros)
  const uint8_t* k_live;   // [batch, ceil(sk/128)] or nullptr: 0 = key block is all padding (skipped);
                           // every batch entry must have at least one live key block
  float* lse;              // [batch, heads, sq] natural-log sum-exp of the masked scores, or nullptr
  // attention dropout (transformer.py:345-346): P is multiplied by keep / (1 - p) AFTER the softmax
  // normalisation — row sums and lse come from the undropped probabilities; element (row, col) =
  // ((b * heads + head) * sq + query, key).  threshold 0 = off.
  DropoutArgs drop;
};

struct AttnBwdArgs {
  uint32_t batch, heads, sq, sk;
  uint32_t causal;
  uint32_t idesc_s;        // M=128, N=64, A/B K-major   (S-type products over 64-row streamed blocks)
  uint32_t idesc_o;        // M=128, N=64, B MN-major    (accumulating products)
  float scale, scale_log2;
  const uint8_t* q_pad;
  const uint8_t* k_pad;
  const uint8_t* q_live;
  const uint8_t* k_live;
  const float* lse;        // [batch, heads, sq] from the forward
  const float* dvec;       // [batch, heads, sq] rowsum(dO * O), from launch_attention_bwd_prep
  DropoutArgs drop;        // the forward's dropout, regenerated: dP = keep / (1 - p) * (dO . V^T)
};

// ---- variable-length forward over token-packed activations (attention_varlen.cu)
struct AttnVarlenItem {      // 32 bytes, read as two int4
  int32_t q_row0;            // first row of the 128-query tile in the packed Q matrix
  int32_t q_valid;           // rows of the tile that belong to the sequence (1..128)
  int32_t k_row0;            // first row of the item's key range in the packed K / V matrices
  int32_t k_len;             // keys in the range (>= 1)
  int32_t head;
  int32_t o_row0;            // output row of the tile's first query
  int32_t lse_idx0;          // index of the first query's entry in the lse buffer
  int32_t reserved;
};
static_assert(sizeof(AttnVarlenItem) == 32, "work items are read as two int4");

struct AttnVarlenArgs {
  const AttnVarlenItem* items;   // device, sorted by decreasing cost; dealt round-robin to the CTAs
  uint32_t n_items;
  uint32_t idesc_s, idesc_o;
  float scale_log2;
  void* out;                     // packed output matrix (row-wise stores of partial tiles)
  int64_t ldo;
  float* lse;                    // optional
  uint32_t sched;                // 0: S(j+1) issued ahead of P.V(j) (production); 1: behind it (A/B, EMDR2_VARLEN_SCHED=1)
};

cudaError_t attention_varlen_prepare();
void launch_attention_varlen(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k, const CUtensorMap& tmap_v,
                             const CUtensorMap& tmap_o, const AttnVarlenArgs& args, bool bf16, int sm_count,
                             cudaStream_t stream);

cudaError_t attention_bwd_prepare();
cudaError_t launch_attention_bwd_prep(bool bf16, const void* dout, int64_t lddo, const void* out, int64_t ldo,
                                      float* dvec, int batch, int heads, int sq, cudaStream_t stream);
// 3-D tensor maps of the backward: 128-row boxes for the tiles a CTA owns / stores, 64-row boxes for
// the operand blocks it streams.
struct AttnBwdMaps {
  CUtensorMap q128, do128, dq128, k64, v64;        // dQ kernel
  CUtensorMap k128, v128, dk128, dv128, q64, do64; // dK/dV kernel
};
cudaError_t launch_attention_bwd(const AttnBwdMaps& maps, const AttnBwdArgs& a, bool bf16, cudaStream_t stream);

cudaError_t attention_prepare();
// Persistent forward (attention_persist.cu): production path; the one-CTA-per-item kernel of
// attention.cu is kept behind EMDR2_ATTN_IMPL=legacy for A/B runs.
cudaError_t attention_persistent_prepare();
void launch_attention_fwd_persistent(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k,
                                     const CUtensorMap& tmap_v, const CUtensorMap& tmap_o,
                                     const AttnArgs& args, bool bf16, int sm_count, cudaStream_t stream);
void launch_attention_fwd(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k,
                          const CUtensorMap& tmap_v, const CUtensorMap& tmap_o,
                          const AttnArgs& args, bool bf16, cudaStream_t stream);

}  // namespace emdr2
