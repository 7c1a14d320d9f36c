#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emdr2 {

constexpr int kMergeThreads = 256;
constexpr int kMergeSortCap = 2048;  // entries sorted per pass in shared memory (12 B each)

// Dense mode: scores/ids are [parts, nq, k]; entries with id < 0 or NaN score are padding.
cudaError_t launch_mips_merge_dense(const float* scores, const int64_t* ids, int parts, int nq, int k,
                                    float* out_scores, int64_t* out_ids, cudaStream_t stream);
// Pool mode: per query q, pool_cnt[q] unordered entries at pool_*[q * pool_cap ...]; the kernel
// resets pool_cnt[q] to 0 when it is done (ready for the next scan on the same stream).
cudaError_t launch_mips_merge_pool(const float* pool_scores, const int64_t* pool_ids,
                                   uint32_t* pool_cnt, uint32_t pool_cap, int nq, int k,
                                   float* out_scores, int64_t* out_ids, cudaStream_t stream);
cudaError_t launch_mips_fill_empty(float* out_scores, int64_t* out_ids, int n, cudaStream_t stream);

}  // namespace emdr2
