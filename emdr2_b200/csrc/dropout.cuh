// Counter-based dropout shared by the attention kernels and the row operators.
//
// The reference applies torch dropout (p = 0.1 by default, megatron/arguments.py:218-221) to the
// attention probabilities (megatron/model/transformer.py:345-346), to every bias-add-residual
// (transformer.py:397-419, 511-515) and to the embedding sum (language_model.py:181), and keeps the
// masks as tensors for the backward pass.  Here a mask is a pure function of (seed, call offset,
// row, column) and is REGENERATED wherever it is needed — forward, both attention backward kernels,
// the hidden-state backward — never stored:
//
//   key(seed, offset)          two 32-bit words, mixed on the host once per launch
//   A(row)  = fmix32(fmix32(key_a ^ row_lo) ^ row_hi ^ key_b) | 1         once per row per thread
//   B(col)  = fmix32(fmix32(seed_lo ^ C ^ col) + seed_hi) | 1             a per-seed table in HBM
//   r       = lo32(A * B)                                                 one IMAD (A, B odd: a bijection of either)
//   keep    = r >= threshold,  threshold = round(p * 2^32)
//
// so an element costs three issue slots (multiply, compare, select) in whichever orientation a
// kernel walks the (row, column) plane — the dK/dV kernel owns KEY rows and walks queries, the
// forward and dQ kernels own QUERY rows and walk keys; a block cipher over 16-element strips
// (Philox) would be cheap in one orientation and 16x the work in the other.  Kept values are scaled
// by 1 / (1 - threshold / 2^32).  Statistical checks: tests/test_dropout.py.
#pragma once
#include <stdint.h>

namespace emdr2 {

struct DropoutArgs {
  uint32_t threshold;        // 0 = dropout off
  uint32_t key_a, key_b;
  float inv_keep;
  const uint32_t* colhash;   // device table B(col), col < the op's column count
};

__host__ __device__ __forceinline__ uint32_t dropout_fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}

__host__ __device__ __forceinline__ uint32_t dropout_row_hash(uint32_t key_a, uint32_t key_b, uint64_t row) {
  return dropout_fmix32(dropout_fmix32(key_a ^ static_cast<uint32_t>(row)) ^ static_cast<uint32_t>(row >> 32) ^ key_b) | 1u;
}

__host__ __device__ __forceinline__ uint32_t dropout_col_hash(uint64_t seed, uint32_t col) {
  return dropout_fmix32(dropout_fmix32((static_cast<uint32_t>(seed) ^ 0x632BE5ABu) ^ col) +
                        static_cast<uint32_t>(seed >> 32)) | 1u;
}

__host__ __device__ __forceinline__ bool dropout_keep(uint32_t a, uint32_t b, uint32_t threshold) {
  return a * b >= threshold;   // the comparison is decided by the product's high bits, which depend on every bit of a and b
}

inline DropoutArgs make_dropout_args(float p, uint64_t seed, uint64_t offset, const uint32_t* colhash) {
  DropoutArgs d{};
  if (!(p > 0.f)) return d;
  double t = static_cast<double>(p) * 4294967296.0 + 0.5;
  if (t > 4294967295.0) t = 4294967295.0;
  d.threshold = static_cast<uint32_t>(t);
  uint32_t a = dropout_fmix32(static_cast<uint32_t>(seed) ^ 0x9E3779B9u);
  a = dropout_fmix32(a ^ static_cast<uint32_t>(seed >> 32));
  a = dropout_fmix32(a ^ static_cast<uint32_t>(offset));
  d.key_a = a;
  d.key_b = dropout_fmix32(a ^ static_cast<uint32_t>(offset >> 32) ^ 0x7F4A7C15u);
  d.inv_keep = static_cast<float>(1.0 / (1.0 - static_cast<double>(d.threshold) / 4294967296.0));
  d.colhash = colhash;
  return d;
}

}  // namespace emdr2
