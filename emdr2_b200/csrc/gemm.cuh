// Dense GEMM with fused epilogue for the BERT/T5 blocks, sm_100a only.
//
//   D[M,N] = epilogue( A[M,K] · B[N,K]ᵀ )      A, B, D 16-bit (fp16 or bf16), fp32 accumulate
//
// i.e. torch.nn.functional.linear(x, W): every projection of the reference's transformer layer
// (megatron/mpu/layers.py:255,353 called from megatron/model/transformer.py:97,107,223,243,262,389)
// and the tied LM head (megatron/model/language_model.py:28-42).  The epilogue fuses what the
// reference runs as separate ATen kernels:
//   +bias[N]                          (ColumnParallelLinear / RowParallelLinear bias add)
//   GeLU(erf)                         (transformer.py:80,99-104, F.gelu default)
//   +residual[M,N]                    (bias_dropout_add with dropout off, transformer.py:397-401)
//
// Structure (one persistent CTA per SM, 128x256 output tiles, N fastest so concurrently running
// CTAs share A rows through L2):
//   warp 0      TMA producer: A box 128x64, B box 256x64 elements per stage, 4-stage mbarrier ring
//   warp 1      tcgen05.mma cta_group::1 kind::f16, M=128 N=256 K=16, accumulators double-buffered
//               in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: tcgen05.ld -> bias/GeLU/residual in fp32 -> 16-bit -> swizzled smem
//               staging -> TMA store (64-column boxes), two column halves in parallel; the residual
//               tile is prefetched by TMA into the same staging buffer one chunk ahead
//   warps 4-19  "wide" variants: four column quarters in parallel — four warps per scheduler hide the GeLU chain's
//               latency; write-only epilogues (bias / GeLU / pre-activation) through 32-column boxes, aux epilogues
//               (residual / GeLU backward) with three operand stages and the aux tile prefetched a tile ahead
// Roofline: tensor pipe (2·M·N·K flop); DRAM traffic is A once + D once (B stays in L2).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace emdr2 {

constexpr int kGemmBM = 128;
constexpr int kGemmBN = 256;
constexpr int kGemmBK = 64;
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 384;
constexpr int kGemmStageA = kGemmBM * kGemmBK * 2;            // 16 KiB
constexpr int kGemmStageB = kGemmBN * kGemmBK * 2;            // 32 KiB
constexpr int kGemmStageBytes = kGemmStageA + kGemmStageB;    // 48 KiB
constexpr int kGemmOutBytes = 2 * kGemmBM * 64 * 2;           // two 128x64 staging boxes
constexpr int kGemmBarBytes = 2048;                              // barriers + the staged bias slice
constexpr int kGemmSmemBytes = kGemmStages * kGemmStageBytes + kGemmOutBytes + kGemmBarBytes + 1024;
constexpr int kGemmWideThreads = 128 + 16 * 32;               // "wide" variants: sixteen epilogue warps (gemm.cu)
constexpr int kGemmAuxStages = 3;                             // wide variant with an aux tile: three operand stages and
constexpr int kGemmAuxOutBytes = 4 * kGemmBM * 64 * 2;        // four 128x64 boxes (aux in, result out, in place)

constexpr uint32_t kGemmBias = 1u;
constexpr uint32_t kGemmGelu = 2u;
constexpr uint32_t kGemmResidual = 4u;
constexpr uint32_t kGemmAccumF32 = 8u;    // D is fp32 [M, ldd32]; results are atomically ADDED (split-K, grad accumulation)
constexpr uint32_t kGemmGeluBwd = 16u;    // D = acc * GeLU'(aux), aux [M, N] 16-bit arrives like the residual
constexpr uint32_t kGemmPreact = 32u;     // also store acc + bias (before GeLU) through tmap_p

struct GemmArgs {
  uint32_t M, N, K;
  uint32_t tiles_m, tiles_n;
  uint32_t idesc;
  uint32_t flags;
  uint32_t splits;         // split-K factor (>1 only with kGemmAccumF32)
  uint32_t kb_per_split;   // 64-wide k blocks per split
  uint32_t ldd32;          // row pitch of out32 (elements)
  const void* bias;        // [N] 16-bit
  float* out32;            // fp32 output for kGemmAccumF32
};

cudaError_t gemm_prepare();
// tmap_r: residual / GeLU-backward aux [M, N] (box 128 x 64), read only with those flags; tmap_p:
// pre-activation output, written only with kGemmPreact; pass tmap_d for the unused ones.
// a_mn / b_mn: operand stored [k, m] / [k, n] row-major (tensor maps with 64 x 64 boxes).
// epi: 0 = eight epilogue warps (any epilogue); 1 = sixteen, write-only epilogues (no aux tile): tmap_d / tmap_p
// must carry 128-row x 32-column boxes with the 64-byte swizzle; 2 = sixteen with an aux tile (residual / GeLU
// backward, no pre-activation output), the usual 64-column boxes.  1 and 2: no fp32 accumulation, splits == 1.
cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const CUtensorMap& tmap_d,
                        const CUtensorMap& tmap_r, const CUtensorMap& tmap_p, const GemmArgs& args,
                        bool bf16, bool a_mn, bool b_mn, int epi, int grid, cudaStream_t stream);

// CTA-pair variant (gemm_pair.cu): K-major operands, 16-bit output, no split-K; 256 x 256 tiles shared by
// two CTAs (tcgen05.mma.cta_group::2).  tmap_b_half: B [N, K] with a 128 x 64 box (each CTA stages half
// of the tile's B rows).  grid must be even (whole pairs).
cudaError_t gemm_pair_prepare();
cudaError_t launch_gemm_pair(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b_half, const CUtensorMap& tmap_d,
                             const CUtensorMap& tmap_r, const CUtensorMap& tmap_p, const GemmArgs& args,
                             bool bf16, int grid, cudaStream_t stream);

}  // namespace emdr2
