// Final ranking of the candidates that survived the scan: (score desc, id asc), best k per query.
//
// Replaces the reference's global selection step — the per-GPU score slabs copied into C[nq, N] and
// the torch.topk over all N columns (megatron/data/emdr2_index.py:284-295).  Here the input is a
// few hundred surviving (score, id) pairs per query (the scan CTAs' pools) or `parts` per-shard
// top-k lists (the payload of the all-gather), so one thread block per query bitonic-sorts them in
// shared memory; inputs longer than one pass are folded chunk by chunk, carrying the best k.
#include "mips_merge.cuh"

namespace emdr2 {
namespace {

__device__ __forceinline__ uint32_t ord32(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unord32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
// a ranks before b ?
__device__ __forceinline__ bool before(uint32_t sa, int64_t ia, uint32_t sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

template <bool kPool>
__global__ void __launch_bounds__(kMergeThreads)
mips_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids,
                  uint32_t* __restrict__ pool_cnt, uint32_t pool_cap, int parts, int nq, int k,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ int64_t sid[kMergeSortCap];
  __shared__ uint32_t ssc[kMergeSortCap];
  const int q = blockIdx.x;
  const int tid = threadIdx.x;
  const int count = kPool ? static_cast<int>(min(pool_cnt[q], pool_cap)) : parts * k;
  const int chunk = kMergeSortCap - k;

  int best_n = 0;  // entries [0, best_n) of the shared arrays hold the best seen so far, sorted
  for (int base = 0; base < count || base == 0; base += chunk) {
    const int take = min(chunk, count - base);
    for (int e = tid; e < take; e += kMergeThreads) {
      size_t src;
      if (kPool) {
        src = static_cast<size_t>(q) * pool_cap + base + e;
      } else {
        const int idx = base + e;
        const int p = idx / k, j = idx - p * k;
        src = (static_cast<size_t>(p) * nq + q) * k + j;
      }
      const int64_t id = ids[src];
      const float s = scores[src];
      const bool live = id >= 0 && s == s;
      sid[best_n + e] = live ? id : INT64_MAX;
      ssc[best_n + e] = live ? ord32(s) : 0u;
    }
    const int total = best_n + max(take, 0);
    int n = 2;
    while (n < total) n <<= 1;
    for (int e = total + tid; e < n; e += kMergeThreads) {
      sid[e] = INT64_MAX;
      ssc[e] = 0u;
    }
    __syncthreads();
    for (int size = 2; size <= n; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (n >> 1); t += kMergeThreads) {
          const int i = 2 * t - (t & (stride - 1));
          const int j = i + stride;
          const bool fwd = (i & size) == 0;
          const uint32_t si = ssc[i], sj = ssc[j];
          const int64_t ii = sid[i], ij = sid[j];
          if (before(sj, ij, si, ii) == fwd) {
            ssc[i] = sj;
            ssc[j] = si;
            sid[i] = ij;
            sid[j] = ii;
          }
        }
        __syncthreads();
      }
    }
    best_n = min(total, k);
    if (take <= 0) break;
  }

  for (int e = tid; e < k; e += kMergeThreads) {
    const size_t dst = static_cast<size_t>(q) * k + e;
    const bool live = e < best_n && sid[e] != INT64_MAX;
    out_scores[dst] = live ? unord32(ssc[e]) : __uint_as_float(0xff800000u);
    out_ids[dst] = live ? sid[e] : -1;
  }
  if (kPool && tid == 0) pool_cnt[q] = 0;
}

__global__ void mips_fill_empty_kernel(float* out_scores, int64_t* out_ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    out_scores[i] = __uint_as_float(0xff800000u);
    out_ids[i] = -1;
  }
}

}  // namespace

cudaError_t launch_mips_merge_dense(const float* scores, const int64_t* ids, int parts, int nq, int k,
                                    float* out_scores, int64_t* out_ids, cudaStream_t stream) {
  if (k >= kMergeSortCap / 2) return cudaErrorInvalidValue;
  mips_merge_kernel<false><<<nq, kMergeThreads, 0, stream>>>(scores, ids, nullptr, 0u, parts, nq, k,
                                                             out_scores, out_ids);
  return cudaGetLastError();
}

cudaError_t launch_mips_merge_pool(const float* pool_scores, const int64_t* pool_ids,
                                   uint32_t* pool_cnt, uint32_t pool_cap, int nq, int k,
                                   float* out_scores, int64_t* out_ids, cudaStream_t stream) {
  if (k >= kMergeSortCap / 2) return cudaErrorInvalidValue;
  mips_merge_kernel<true><<<nq, kMergeThreads, 0, stream>>>(pool_scores, pool_ids, pool_cnt, pool_cap,
                                                            0, nq, k, out_scores, out_ids);
  return cudaGetLastError();
}

cudaError_t launch_mips_fill_empty(float* out_scores, int64_t* out_ids, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  mips_fill_empty_kernel<<<(n + 255) / 256, 256, 0, stream>>>(out_scores, out_ids, n);
  return cudaGetLastError();
}

}  // namespace emdr2
