// Backward of the row-wise operators (see rowops.cuh): LayerNorm, bias (column sums), the token
// log-probability / cross-entropy, and the embedding sum.  fp32 math; parameter gradients are
// ACCUMULATED into fp32 buffers with atomics (main-grad style), activation gradients are 16-bit.
#include "rowops.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace emdr2 {
namespace {

constexpr int kMaxChunks = 4;

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

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = dy * gamma, xhat = (x - mean) * rstd
// dgamma += sum_rows dy * xhat, dbeta += sum_rows dy.
// One warp per row, rows handed out with a grid stride over RESIDENT blocks (two per SM).  Register budget is what
// bounds this kernel: the per-thread dgamma / dbeta accumulators are 2 x 8 floats per 256-column chunk, so the kernel
// is instantiated for the number of chunks the width needs (three for h = 768), keeps dy / x / dres of the row as the
// raw 16-byte words across the two warp reductions and unpacks them again for dx, and keeps gamma packed — 120 instead
// of 213 registers, two blocks per SM instead of one (measured 2.0 -> see profiles/README.md TB/s at 185 600 x 768).
template <bool kBf16, int kChunks>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const uint16_t* __restrict__ dy, int64_t ldy, const uint16_t* __restrict__ x, int64_t ldx,
                     const uint16_t* __restrict__ gamma, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const uint16_t* __restrict__ dres, int64_t ldr,
                     uint16_t* __restrict__ dx, int64_t lddx, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int rows, int h) {
  __shared__ float s_red[8][kChunks * 256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc_g[kChunks][8], acc_b[kChunks][8];
  uint4 gam_raw[kChunks];
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    const int col = (c * 32 + lane) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc_g[c][i] = acc_b[c][i] = 0.f;
    gam_raw[c] = col < h ? __ldg(reinterpret_cast<const uint4*>(gamma + col)) : make_uint4(0u, 0u, 0u, 0u);
  }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    uint4 d_raw[kChunks], x_raw[kChunks], r_raw[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {     // all of the row's loads first
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
        d_raw[c] = __ldg(reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * ldy + col));
        x_raw[c] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx + col));
        if (dres) r_raw[c] = __ldg(reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * ldr + col));
      }
    }
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
        float d[8], xv[8], gm[8];
        unpack8<kBf16>(d_raw[c], d);
        unpack8<kBf16>(x_raw[c], xv);
        unpack8<kBf16>(gam_raw[c], gm);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xv[i] - mu) * rs;
          const float g = d[i] * gm[i];
          s1 += g;
          s2 = fmaf(g, xh, s2);
          acc_g[c][i] = fmaf(d[i], xh, acc_g[c][i]);
          acc_b[c][i] += d[i];
        }
      }
    }
    const float c1 = warp_sum(s1) / static_cast<float>(h);
    const float c2 = warp_sum(s2) / static_cast<float>(h);
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
        float d[8], xv[8], gm[8], o[8];
        unpack8<kBf16>(d_raw[c], d);
        unpack8<kBf16>(x_raw[c], xv);
        unpack8<kBf16>(gam_raw[c], gm);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float xh = (xv[i] - mu) * rs;       // the same expressions as above: bit-identical g and xhat
          const float g = d[i] * gm[i];
          o[i] = rs * (g - c1 - xh * c2);
        }
        if (dres) {
          float r[8];
          unpack8<kBf16>(r_raw[c], r);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] += r[i];
        }
        *reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * lddx + col) = pack8<kBf16>(o);
      }
    }
  }
  // block reduction of the parameter gradients, then one atomic per column per block
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < h) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s_red[warp][col + i] = pass == 0 ? acc_g[c][i] : acc_b[c][i];
      }
    }
    __syncthreads();
    float* dst = pass == 0 ? dgamma : dbeta;
    if (dst) {
      for (int col = threadIdx.x; col < h; col += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w][col];
        atomicAdd(dst + col, t);
      }
    }
  }
}

// out[n] += sum_m dy[m, n]   (bias gradient).  Block = 32 column groups (8 columns = 16 B each) x 8
// row lanes; every thread keeps 4 independent 16-byte loads in flight, the 8 row lanes are folded
// through shared memory and each block issues one atomic per column.
template <bool kBf16>
__global__ void __launch_bounds__(256)
colsum_kernel(const uint16_t* __restrict__ dy, int64_t ld, float* __restrict__ out, int rows, int n) {
  __shared__ float s_part[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < n) {
    const int stride = gridDim.y * 8;
    int r = blockIdx.y * 8 + ty;
    for (; r + 3 * stride < rows; r += 4 * stride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg(reinterpret_cast<const uint4*>(dy + static_cast<size_t>(r + u * stride) * ld + col));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8<kBf16>(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      }
    }
    for (; r < rows; r += stride) {
      float f[8];
      unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dy + static_cast<size_t>(r) * ld + col)), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_part[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;           // 256 threads -> 256 columns of this block
  const int gc = blockIdx.x * 256 + c;
  if (gc < n) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_part[w][c];
    atomicAdd(out + gc, t);
  }
}

// dlogits[r, v] = g[r] * (1[v == label_r] - exp(logits[r, v] - lse[r])): backward of token_logprob.
template <bool kBf16>
__global__ void __launch_bounds__(256)
token_logprob_bwd_kernel(const uint16_t* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                         const float* __restrict__ lse, const float* __restrict__ g,
                         uint16_t* __restrict__ dlogits, int64_t ldd, int vocab) {
  const int row = blockIdx.x;
  const float gr = g[row], l = lse[row];
  const int64_t lab = labels[row];
  const bool lab_ok = lab >= 0 && lab < vocab;
  const uint16_t* lr = logits + static_cast<size_t>(row) * ld;
  uint16_t* dr = dlogits + static_cast<size_t>(row) * ldd;
  for (int col = threadIdx.x * 8; col < vocab; col += 256 * 8) {
    float f[8], o[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(lr + col)), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float onehot = (lab_ok && lab == col + i) ? 1.f : 0.f;
      o[i] = lab_ok ? gr * (onehot - __expf(f[i] - l)) : 0.f;
    }
    *reinterpret_cast<uint4*>(dr + col) = pack8<kBf16>(o);
  }
}

// dword[ids[t]] += dx[t]  (scattered fp32 atomics: rows of different tokens rarely collide)
template <bool kBf16>
__global__ void __launch_bounds__(256)
embedding_bwd_word_kernel(const uint16_t* __restrict__ dx, const int64_t* __restrict__ ids,
                          float* __restrict__ dword, int tokens, int h, int vocab) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= tokens) return;
  int64_t id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  for (int col = lane * 8; col < h; col += 256) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(dx + static_cast<size_t>(t) * h + col));
    if ((raw.x | raw.y | raw.z | raw.w) == 0u) continue;     // padding positions carry exact zeros
    float f[8];
    unpack8<kBf16>(raw, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(dword + static_cast<size_t>(id) * h + col + i, f[i]);
  }
}

// dpos[p] += sum_b dx[b * seq + p];  dtype_emb[ty] += sum over tokens of that type.  One thread per
// (position, 8 columns) walks the batch, so the hot rows (a position is shared by every sequence,
// a token type by half of all tokens) see one atomic per thread instead of one per token.
template <bool kBf16>
__global__ void __launch_bounds__(128)
embedding_bwd_pos_kernel(const uint16_t* __restrict__ dx, const int64_t* __restrict__ types,
                         float* __restrict__ dpos, float* __restrict__ dtype_emb, int tokens, int seq, int h,
                         int num_types) {
  const int pos = blockIdx.x;
  const int col = (blockIdx.y * 128 + threadIdx.x) * 8;
  if (col >= h) return;
  const int batch = tokens / seq;
  float acc[4][8];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  for (int b = 0; b < batch; ++b) {
    const size_t tok = static_cast<size_t>(b) * seq + pos;
    float f[8];
    unpack8<kBf16>(__ldg(reinterpret_cast<const uint4*>(dx + tok * h + col)), f);
    int ty = 0;
    if (types && dtype_emb) {
      const int64_t raw = types[tok];
      ty = raw < 0 ? 0 : (raw >= num_types ? num_types - 1 : static_cast<int>(raw));
      ty = ty > 3 ? 3 : ty;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (t == ty) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] += f[i];
      }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float total = acc[0][i] + acc[1][i] + acc[2][i] + acc[3][i];
    if (dpos) atomicAdd(dpos + static_cast<size_t>(pos) * h + col + i, total);
    if (types && dtype_emb) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < num_types && acc[t][i] != 0.f) atomicAdd(dtype_emb + static_cast<size_t>(t) * h + col + i, acc[t][i]);
    }
  }
}

}  // namespace

cudaError_t launch_layernorm_bwd(bool bf16, const void* dy, int64_t ldy, const void* x, int64_t ldx,
                                 const void* gamma, const float* mean, const float* rstd, const void* dres,
                                 int64_t ldr, void* dx, int64_t lddx, float* dgamma, float* dbeta, int rows,
                                 int h, int sm_count, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (h % 8 || h > kMaxChunks * 256) return cudaErrorInvalidValue;
  int grid = (rows + 7) / 8;
  if (grid > sm_count * 2) grid = sm_count * 2;      // resident blocks (two per SM): one parameter-gradient fold each
  auto p = [](const void* v) { return static_cast<const uint16_t*>(v); };
  const int chunks = (h + 255) / 256;
#define EMDR2_LN_BWD(BF, CH)                                                                                       \
  layernorm_bwd_kernel<BF, CH><<<grid, 256, 0, stream>>>(p(dy), ldy, p(x), ldx, p(gamma), mean, rstd, p(dres), ldr, \
                                                         static_cast<uint16_t*>(dx), lddx, dgamma, dbeta, rows, h)
  if (bf16) {
    if (chunks <= 1) EMDR2_LN_BWD(true, 1);
    else if (chunks == 2) EMDR2_LN_BWD(true, 2);
    else if (chunks == 3) EMDR2_LN_BWD(true, 3);
    else EMDR2_LN_BWD(true, 4);
  } else {
    if (chunks <= 1) EMDR2_LN_BWD(false, 1);
    else if (chunks == 2) EMDR2_LN_BWD(false, 2);
    else if (chunks == 3) EMDR2_LN_BWD(false, 3);
    else EMDR2_LN_BWD(false, 4);
  }
#undef EMDR2_LN_BWD
  return cudaGetLastError();
}

cudaError_t launch_colsum(bool bf16, const void* dy, int64_t ld, float* out, int rows, int n,
                          cudaStream_t stream) {
  if (rows <= 0 || n <= 0) return cudaSuccess;
  if (n % 8) return cudaErrorInvalidValue;
  const int gx = (n + 255) / 256;
  int gy = (rows + 63) / 64;                      // >= 8 rows per row lane
  const int want = (148 * 8 + gx - 1) / gx;       // ~8 blocks per SM over the whole grid
  if (gy > want) gy = want;
  if (gy < 1) gy = 1;
  dim3 grid(gx, gy);
  if (bf16) colsum_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(dy), ld, out, rows, n);
  else colsum_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(dy), ld, out, rows, n);
  return cudaGetLastError();
}

cudaError_t launch_token_logprob_bwd(bool bf16, const void* logits, int64_t ld, const int64_t* labels,
                                     const float* lse, const float* g, void* dlogits, int64_t ldd, int rows,
                                     int vocab, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (vocab % 8) return cudaErrorInvalidValue;
  if (bf16)
    token_logprob_bwd_kernel<true><<<rows, 256, 0, stream>>>(static_cast<const uint16_t*>(logits), ld, labels, lse,
                                                             g, static_cast<uint16_t*>(dlogits), ldd, vocab);
  else
    token_logprob_bwd_kernel<false><<<rows, 256, 0, stream>>>(static_cast<const uint16_t*>(logits), ld, labels, lse,
                                                              g, static_cast<uint16_t*>(dlogits), ldd, vocab);
  return cudaGetLastError();
}

cudaError_t launch_embedding_bwd(bool bf16, const void* dx, const int64_t* ids, const int64_t* types,
                                 float* dword, float* dpos, float* dtype_emb, int tokens, int seq, int h,
                                 int vocab, int num_types, cudaStream_t stream) {
  if (tokens <= 0) return cudaSuccess;
  if (h % 8 || tokens % seq || num_types > 4) return cudaErrorInvalidValue;
  auto d = static_cast<const uint16_t*>(dx);
  if (dword) {
    const int grid = (tokens + 7) / 8;
    if (bf16) embedding_bwd_word_kernel<true><<<grid, 256, 0, stream>>>(d, ids, dword, tokens, h, vocab);
    else embedding_bwd_word_kernel<false><<<grid, 256, 0, stream>>>(d, ids, dword, tokens, h, vocab);
  }
  if (dpos || (types && dtype_emb)) {
    dim3 grid(seq, (h / 8 + 127) / 128);
    if (bf16)
      embedding_bwd_pos_kernel<true><<<grid, 128, 0, stream>>>(d, types, dpos, dtype_emb, tokens, seq, h, num_types);
    else
      embedding_bwd_pos_kernel<false><<<grid, 128, 0, stream>>>(d, types, dpos, dtype_emb, tokens, seq, h, num_types);
  }
  return cudaGetLastError();
}

}  // namespace emdr2
