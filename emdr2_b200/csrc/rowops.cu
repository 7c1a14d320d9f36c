// See rowops.cuh. One warp per row, 16-byte vectorised loads/stores, fp32 math.
#include "rowops.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace emdr2 {
namespace {

constexpr int kMaxChunks = 4;  // 8 elements per lane per chunk -> h <= 1024

template <bool kBf16>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 p;
    if constexpr (kBf16) p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
    else p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = p.x;
    f[2 * i + 1] = p.y;
  }
}
template <bool kBf16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if constexpr (kBf16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per row, rows handed out with a grid stride over resident blocks: gamma and beta are staged ONCE per block
// in shared memory (they used to be fetched from global memory per row, behind the two reductions).
template <bool kBf16>
__global__ void __launch_bounds__(256, 4)
layernorm_fwd_kernel(const uint16_t* __restrict__ x, int64_t ldx, const uint16_t* __restrict__ gamma,
                     const uint16_t* __restrict__ beta, uint16_t* __restrict__ y, int64_t ldy,
                     int rows, int h, float eps, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out) {
  __shared__ uint4 s_gamma[kMaxChunks * 32], s_beta[kMaxChunks * 32];
  for (int i = threadIdx.x; i < kMaxChunks * 32; i += blockDim.x) {
    if (i * 8 < h) {
      s_gamma[i] = __ldg(reinterpret_cast<const uint4*>(gamma + i * 8));
      s_beta[i] = __ldg(reinterpret_cast<const uint4*>(beta + i * 8));
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int stride = gridDim.x * warps_per_block;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += stride) {
    const uint16_t* xr = x + static_cast<size_t>(row) * ldx;
    float v[kMaxChunks][8];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
        unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(xr + col)), v[c]);
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[c][i];
      }
    }
    const float mean = warp_sum(sum) / static_cast<float>(h);
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float d = v[c][i] - mean;
          sq = fmaf(d, d, sq);
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(h) + eps);
    uint16_t* yr = y + static_cast<size_t>(row) * ldy;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
        float g[8], bt[8], o[8];
        unpack8<kBf16>(s_gamma[c * 32 + lane], g);
        unpack8<kBf16>(s_beta[c * 32 + lane], bt);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf((v[c][i] - mean) * rstd, g[i], bt[i]);
        *reinterpret_cast<uint4*>(yr + col) = pack8<kBf16>(o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

template <bool kBf16>
__global__ void __launch_bounds__(256)
embedding_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ types,
                     const uint16_t* __restrict__ word, const uint16_t* __restrict__ pos,
                     const uint16_t* __restrict__ type_emb, uint16_t* __restrict__ out, int tokens,
                     int seq, int h, int vocab, int num_types, const int32_t* __restrict__ pos_ids, int max_pos) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens) return;
  int64_t id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);   // clamp: never read outside the table
  const uint16_t* wr = word + static_cast<size_t>(id) * h;
  int p = pos_ids ? pos_ids[t] : t % seq;             // packed sequences carry their positions explicitly
  p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
  const uint16_t* pr = pos + static_cast<size_t>(p) * h;
  const uint16_t* tr = nullptr;
  if (types && type_emb) {
    int64_t ty = types[t];
    ty = ty < 0 ? 0 : (ty >= num_types ? num_types - 1 : ty);
    tr = type_emb + static_cast<size_t>(ty) * h;
  }
  uint16_t* o = out + static_cast<size_t>(t) * h;
  for (int col = lane * 8; col < h; col += 256) {
    float a[8], b[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(wr + col)), a);
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(pr + col)), b);
    // the reference adds in the 16-bit dtype: (word + pos) rounded, then + type rounded
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = a[i] + b[i];
    if (tr) {
      uint4 r = pack8<kBf16>(s);
      unpack8<kBf16>(r, s);
      float c[8];
      unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(tr + col)), c);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += c[i];
    }
    *reinterpret_cast<uint4*>(o + col) = pack8<kBf16>(s);
  }
}

template <bool kBf16>
__global__ void __launch_bounds__(256)
token_logprob_kernel(const uint16_t* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                     float* __restrict__ logprob, float* __restrict__ lse_out, int vocab) {
  __shared__ float s_max[8], s_sum[8];
  const int row = blockIdx.x;
  const uint16_t* lr = logits + static_cast<size_t>(row) * ld;
  // online (max, sum exp) per thread over 16-byte chunks
  float m = -INFINITY, s = 0.f;
  const bool vec = (vocab % 8 == 0) && (ld % 8 == 0);
  if (!vec) {  // odd vocabulary sizes (toy configurations): element-wise loads
    for (int col = threadIdx.x; col < vocab; col += 256) {
      const uint16_t raw = lr[col];
      float x;
      if constexpr (kBf16) x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&raw));
      else x = __half2float(*reinterpret_cast<const __half*>(&raw));
      const float nm = fmaxf(m, x);
      s = s * __expf(m - nm) + __expf(x - nm);
      m = nm;
    }
  }
  for (int col = threadIdx.x * 8; vec && col < vocab; col += 256 * 8) {
    float f[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(lr + col)), f);
    float cm = f[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) cm = fmaxf(cm, f[i]);
    const float nm = fmaxf(m, cm);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __expf(f[i] - nm);
    s = s * __expf(m - nm) + acc;
    m = nm;
  }
  // warp, then block combine of (m, s)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const float os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = (m == -INFINITY ? 0.f : s * __expf(m - nm)) + (om == -INFINITY ? 0.f : os * __expf(om - nm));
    m = nm;
  }
  if ((threadIdx.x & 31) == 0) {
    s_max[threadIdx.x >> 5] = m;
    s_sum[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = s_max[0], bs = s_sum[0];
    for (int w = 1; w < 8; ++w) {
      const float om = s_max[w], os = s_sum[w];
      const float nm = fmaxf(bm, om);
      bs = (bm == -INFINITY ? 0.f : bs * __expf(bm - nm)) + (om == -INFINITY ? 0.f : os * __expf(om - nm));
      bm = nm;
    }
    const float lse = bm + logf(bs);
    if (lse_out) lse_out[row] = lse;
    const int64_t lab = labels[row];
    float lp = 0.f;
    if (lab >= 0 && lab < vocab) {
      const uint16_t raw = lr[lab];
      float x;
      if constexpr (kBf16) x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&raw));
      else x = __half2float(*reinterpret_cast<const __half*>(&raw));
      lp = x - lse;
    }
    logprob[row] = lp;
  }
}

}  // namespace

cudaError_t launch_token_logprob(bool bf16, const void* logits, int64_t ld, const int64_t* labels,
                                 float* logprob, float* lse, int rows, int vocab,
                                 cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  auto ls = static_cast<const uint16_t*>(logits);
  if (bf16) token_logprob_kernel<true><<<rows, 256, 0, stream>>>(ls, ld, labels, logprob, lse, vocab);
  else token_logprob_kernel<false><<<rows, 256, 0, stream>>>(ls, ld, labels, logprob, lse, vocab);
  return cudaGetLastError();
}

cudaError_t launch_layernorm_fwd(bool bf16, const void* x, int64_t ldx, const void* gamma,
                                 const void* beta, void* y, int64_t ldy, int rows, int h, float eps,
                                 float* mean, float* rstd, int sm_count, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (h % 8 || h > kMaxChunks * 256) return cudaErrorInvalidValue;
  const int warps = 8;
  int grid = (rows + warps - 1) / warps;
  if (sm_count > 0 && grid > sm_count * 8) grid = sm_count * 8;   // two waves of resident blocks; warps stride over the rows
  auto xs = static_cast<const uint16_t*>(x);
  auto gs = static_cast<const uint16_t*>(gamma);
  auto bs = static_cast<const uint16_t*>(beta);
  auto ys = static_cast<uint16_t*>(y);
  if (bf16)
    layernorm_fwd_kernel<true><<<grid, warps * 32, 0, stream>>>(xs, ldx, gs, bs, ys, ldy, rows, h, eps, mean, rstd);
  else
    layernorm_fwd_kernel<false><<<grid, warps * 32, 0, stream>>>(xs, ldx, gs, bs, ys, ldy, rows, h, eps, mean, rstd);
  return cudaGetLastError();
}

cudaError_t launch_embedding_fwd(bool bf16, const int64_t* ids, const int64_t* types,
                                 const void* word, const void* pos, const void* type_emb, void* out,
                                 int tokens, int seq, int h, int vocab, int num_types,
                                 cudaStream_t stream, const int32_t* pos_ids, int max_pos) {
  if (tokens <= 0) return cudaSuccess;
  if (h % 8) return cudaErrorInvalidValue;
  if (max_pos <= 0) max_pos = seq;
  const int warps = 8;
  const int grid = (tokens + warps - 1) / warps;
  auto ws = static_cast<const uint16_t*>(word);
  auto ps = static_cast<const uint16_t*>(pos);
  auto ts = static_cast<const uint16_t*>(type_emb);
  auto os = static_cast<uint16_t*>(out);
  if (bf16)
    embedding_fwd_kernel<true><<<grid, warps * 32, 0, stream>>>(ids, types, ws, ps, ts, os, tokens, seq, h, vocab, num_types, pos_ids, max_pos);
  else
    embedding_fwd_kernel<false><<<grid, warps * 32, 0, stream>>>(ids, types, ws, ps, ts, os, tokens, seq, h, vocab, num_types, pos_ids, max_pos);
  return cudaGetLastError();
}

}  // namespace emdr2

// ------------------------------------------------------------------------------------- dropout
namespace emdr2 {
namespace {

__global__ void dropout_colhash_kernel(uint64_t seed, uint32_t* __restrict__ table, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) table[c] = dropout_col_hash(seed, static_cast<uint32_t>(c));
}

__global__ void dropout_mask_kernel(DropoutArgs d, uint8_t* __restrict__ mask, int64_t rows, int cols) {
  const int64_t r = blockIdx.x;
  const uint32_t a = dropout_row_hash(d.key_a, d.key_b, static_cast<uint64_t>(r));
  for (int c = threadIdx.x; c < cols; c += blockDim.x)
    mask[r * cols + c] = dropout_keep(a, __ldg(d.colhash + c), d.threshold) ? 1 : 0;
}

// One warp per row, 8 elements per lane per step (16-byte loads/stores).
template <bool kBf16>
__global__ void __launch_bounds__(256)
dropout_add_kernel(const uint16_t* __restrict__ y, int64_t ldy, const uint16_t* __restrict__ res, int64_t ldr,
                   uint16_t* __restrict__ out, int64_t ldo, int rows, int cols, DropoutArgs d) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int row = blockIdx.x * warps + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps) {
    const uint32_t a = dropout_row_hash(d.key_a, d.key_b, static_cast<uint64_t>(row));
    const uint16_t* yr = y + static_cast<size_t>(row) * ldy;
    const uint16_t* rr = res ? res + static_cast<size_t>(row) * ldr : nullptr;
    uint16_t* orow = out + static_cast<size_t>(row) * ldo;
    for (int col = lane * 8; col < cols; col += 256) {
      float v[8], r[8];
      unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(yr + col)), v);
      const uint4 h0 = __ldg(reinterpret_cast<const uint4*>(d.colhash + col));
      const uint4 h1 = __ldg(reinterpret_cast<const uint4*>(d.colhash + col + 4));
      const uint32_t hb[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
      if (rr) unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(rr + col)), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float kept = dropout_keep(a, hb[i], d.threshold) ? v[i] * d.inv_keep : 0.f;
        v[i] = rr ? r[i] + kept : kept;
      }
      *reinterpret_cast<uint4*>(orow + col) = pack8<kBf16>(v);
    }
  }
}

}  // namespace

cudaError_t launch_dropout_colhash(uint64_t seed, uint32_t* table, int n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  dropout_colhash_kernel<<<(n + 255) / 256, 256, 0, stream>>>(seed, table, n);
  return cudaGetLastError();
}

cudaError_t launch_dropout_mask(const DropoutArgs& d, uint8_t* mask, int64_t rows, int cols, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  dropout_mask_kernel<<<static_cast<unsigned>(rows), 256, 0, stream>>>(d, mask, rows, cols);
  return cudaGetLastError();
}

cudaError_t launch_dropout_add(bool bf16, const void* y, int64_t ldy, const void* residual, int64_t ldr, void* out,
                               int64_t ldo, int rows, int cols, const DropoutArgs& d, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const int warps = 8;
  const int blocks = static_cast<int>(((rows + warps - 1) / warps) < 148 * 16 ? (rows + warps - 1) / warps : 148 * 16);
  if (bf16)
    dropout_add_kernel<true><<<blocks, 32 * warps, 0, stream>>>(
        static_cast<const uint16_t*>(y), ldy, static_cast<const uint16_t*>(residual), ldr,
        static_cast<uint16_t*>(out), ldo, rows, cols, d);
  else
    dropout_add_kernel<false><<<blocks, 32 * warps, 0, stream>>>(
        static_cast<const uint16_t*>(y), ldy, static_cast<const uint16_t*>(residual), ldr,
        static_cast<uint16_t*>(out), ldo, rows, cols, d);
  return cudaGetLastError();
}

}  // namespace emdr2
