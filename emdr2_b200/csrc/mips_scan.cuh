// Fused brute-force MIPS scan for sm_100a: tcgen05 GEMM over TMA-staged evidence tiles with an
// in-kernel top-k, so per-passage scores never leave the SM.
//
// Replaces DistributedBruteForceIndex.search_mips_index (reference megatron/data/emdr2_index.py:
// 268-305: torch.matmul :281 -> C[nq,N] fp16 :284-292 -> torch.topk :295) for one evidence shard.
//
// Shape of the computation (one persistent CTA per SM, tiles of 128 evidence rows, round-robin):
//   D[64 queries, 128 rows] (fp32, TMEM) = Q[64, d] (smem resident, A operand) x E_tile[128, d]^T
//   (B operand, streamed by TMA in 64-column K blocks through an mbarrier ring).
// MMA M = 64 queries puts query m on TMEM lane (m % 16) + 32 * (m / 16), so epilogue warp w owns
// queries 16w..16w+15 on its lanes 0..15 and every query's top-k state is private to one thread:
// a candidate list in shared memory, a running threshold in registers, no atomics.
//
// Threshold filtering: a score is appended to its query's candidate list only if it can still be
// in the top k.  Two sources bound it from below:
//   (1) local  — the k-th best (score,row) this CTA has seen for the query (after a compaction);
//   (2) shared — every CTA publishes its running per-query maximum; the k-th largest of those
//       per-CTA maxima is a score that at least k distinct rows reach, so anything strictly below
//       it cannot be in the top k.  One helper warp per CTA recomputes it for one query and
//       publishes it with an atomicMax; epilogue threads pick it up once per tile.  Pure
//       optimisation: results do not depend on its timing, only the amount of work does.
// With 148 CTAs and k = 50 the shared bound sits at roughly the 0.4/n quantile after n rows per CTA,
// so after the first tile almost nothing passes and compactions (a warp bitonic sort) are rare.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace emdr2 {

constexpr int kQ = 64;                               // queries per launch == MMA M
constexpr int kTileN = 128;                          // evidence rows per tile == MMA N
constexpr int kBlockK = 64;                          // elements per K block (128 B rows, SW128)
constexpr int kUmmaK = 16;                           // K per tcgen05.mma (16-bit inputs)
constexpr int kStageBytes = kTileN * kBlockK * 2;    // 16 KiB per pipeline stage
constexpr int kQBlockBytes = kQ * kBlockK * 2;       // 8 KiB per resident query K block
constexpr int kCap = 96;                             // candidate slots per query
constexpr int kCandStride = 98;                      // entries; 98*8 B is 16-B aligned
constexpr int kCandBytes = kQ * kCandStride * 8;     // 50176 B
constexpr int kMaxStages = 12;
constexpr int kMaxK = 64;                            // k + 32 <= kCap
constexpr int kScanThreads = 256;                    // warps 0-3: TMA, MMA, service, spare; 4-7 epilogue
constexpr int kTmemCols = 2 * kTileN;                // double-buffered accumulator
constexpr int kMaxCtas = 256;                        // rows of the shared-max table
constexpr int kBarBytes = 1024;
constexpr int kPoolCap = kMaxCtas * kMaxK;          // worst case: every CTA keeps k rows of a query

constexpr uint32_t kFlagShare = 1u;
constexpr uint32_t kFlagProbe = 2u;

struct ScanArgs {
  uint32_t n_rows;
  uint32_t nq;
  uint32_t k;
  uint32_t num_kb;
  uint32_t num_stages;
  uint32_t num_tiles;
  uint32_t idesc;
  uint32_t epoch;
  uint32_t flags;
  uint32_t probe_timeout_ns;
  const int64_t* ids;   // [n_rows] or nullptr
  int64_t id_base;
  float* pool_scores;   // [kQ, pool_cap] surviving candidates of all CTAs, per query
  int64_t* pool_ids;    // [kQ, pool_cap]
  uint32_t* pool_cnt;   // [kQ] fill level (reset to 0 by the merge kernel)
  uint32_t pool_cap;
  uint64_t* gmax;       // [kQ, kMaxCtas] (epoch << 32 | ordered score)
  uint64_t* gthr;       // [kQ]           (epoch << 32 | ordered score)
  unsigned long long* stats;  // [4]: appends, compactions, max probe wait ns, sum of per-warp waits
};

// Dynamic shared memory layout (offsets from a 1024-B aligned base).
struct ScanSmemLayout {
  uint32_t off_q, off_stage, off_cand, off_bar, total;
};
inline ScanSmemLayout scan_smem_layout(uint32_t num_kb, uint32_t num_stages) {
  ScanSmemLayout l;
  l.off_q = 0;
  l.off_stage = num_kb * kQBlockBytes;
  l.off_cand = l.off_stage + num_stages * kStageBytes;
  l.off_bar = l.off_cand + kCandBytes;
  l.total = l.off_bar + kBarBytes + 1024;  // + slack for manual 1024-B alignment
  return l;
}

void launch_mips_scan(const CUtensorMap& tmap_q, const CUtensorMap& tmap_e, const ScanArgs& args,
                      int grid, uint32_t smem_bytes, cudaStream_t stream);
cudaError_t mips_scan_prepare(uint32_t max_smem_bytes);

}  // namespace emdr2
