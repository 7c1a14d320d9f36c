// Backward of the fused attention (attention.cu), sm_100a only, head dim 64.
//
// Given dO, the saved Q, K, V, O and the per-row log-sum-exp L of the forward:
//   D_i   = sum_d dO_i,d * O_i,d
//   P_ij  = exp(s_ij - L_i),  s_ij = scale * q_i.k_j with masked entries replaced by -10000
//   dV_j  = sum_i P_ij dO_i
//   dP_ij = dO_i . v_j
//   dS_ij = P_ij (dP_ij - D_i) * scale, and 0 where the score was masked (masked_fill replaces the
//           score, so no gradient reaches q.k there — reference megatron/model/bert_model.py:31-33,
//           t5_model.py:28-30 under autograd)
//   dQ_i  = sum_j dS_ij k_j,   dK_j = sum_i dS_ij q_i
// This is the autograd of transformer.py:301-383 (baddbmm, mask+softmax, bmm) without ever
// materialising the [b, np, sq, sk] tensors.  Two kernels, both shaped like the forward (TMA ring,
// tcgen05 S-type products into TMEM, one thread per row for the elementwise part, P/dS handed back
// to the tensor core through shared memory), so neither needs atomics:
//   attention_bwd_dq_kernel    one CTA per 128-query block, loops over key blocks,   dQ in TMEM
//   attention_bwd_dkv_kernel   one CTA per 128-key block,   loops over query blocks, dK and dV in TMEM
#include "attention.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "gelu.cuh"   // packed fp32 helpers (f2_fma / f2_mul: two IEEE fp32 operations per issued instruction)
#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr float kMaskedLog2 = -10000.0f * 1.4426950408889634f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kTile = kAttnTileBytes;   // 128 x 64 x 2 B

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool kBf16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (kBf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (kBf16) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
  else return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

// D[b, h, i] = sum_d dO[b, i, h, d] * O[b, i, h, d]
// Eight lanes per (token, head): each loads ONE 16-byte chunk of the head's 128-byte rows of dO and O (a warp
// instruction covers four whole rows: fully coalesced) and the eight partial sums are folded by shuffles; four
// (token, head) units per thread keep eight independent loads in flight.  (One thread per unit with eight strided
// 16-byte loads ran at half of the copy bandwidth, latency-bound.)
constexpr int kPrepUnroll = 4;
template <bool kBf16>
__global__ void __launch_bounds__(256)
attention_bwd_prep_kernel(const uint16_t* __restrict__ dout, int64_t lddo, const uint16_t* __restrict__ out,
                          int64_t ldo, float* __restrict__ dvec, int batch, int heads, int sq) {
  const uint32_t total = static_cast<uint32_t>(batch) * sq * heads;            // (token, head) units, < 2^31 (checked on the host)
  const uint32_t chunk = threadIdx.x & 7u;
  const uint32_t unit0 = (blockIdx.x * 256u + threadIdx.x) >> 3;
  const uint32_t stride = (gridDim.x * 256u) >> 3;
  const uint32_t uheads = static_cast<uint32_t>(heads), usq = static_cast<uint32_t>(sq);
  uint4 x[kPrepUnroll], y[kPrepUnroll];
  uint32_t tok[kPrepUnroll], hd[kPrepUnroll];
#pragma unroll
  for (int j = 0; j < kPrepUnroll; ++j) {
    const uint32_t u = unit0 + j * stride;
    tok[j] = u / uheads;
    hd[j] = u - tok[j] * uheads;
    if (u < total) {
      x[j] = __ldg(reinterpret_cast<const uint4*>(dout + static_cast<int64_t>(tok[j]) * lddo + hd[j] * 64 + chunk * 8));
      y[j] = __ldg(reinterpret_cast<const uint4*>(out + static_cast<int64_t>(tok[j]) * ldo + hd[j] * 64 + chunk * 8));
    } else {
      x[j] = y[j] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
#pragma unroll
  for (int j = 0; j < kPrepUnroll; ++j) {
    const uint32_t xw[4] = {x[j].x, x[j].y, x[j].z, x[j].w}, yw[4] = {y[j].x, y[j].y, y[j].z, y[j].w};
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 p = unpack2<kBf16>(xw[i]), q = unpack2<kBf16>(yw[i]);
      acc = fmaf(p.x, q.x, acc);
      acc = fmaf(p.y, q.y, acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (chunk == 0 && unit0 + j * stride < total) {
      const uint32_t bi = tok[j] / usq, i = tok[j] - bi * usq;
      dvec[(static_cast<int64_t>(bi) * heads + hd[j]) * sq + i] = acc;
    }
  }
}

struct BwdBars {
  uint64_t fixed_full;     // the CTA's own tiles (Q,dO resp. K,V)
  uint64_t ring_full[2];   // streamed 64-row tiles (K,V resp. Q,dO)
  uint64_t ring_empty[2];
  uint64_t sdp_full;       // S and dP products landed in TMEM
  uint64_t sdp_empty;      // row threads have read them (4 warps)
  uint64_t pds_full;       // P / dS written to shared memory (4 warps)
  uint64_t acc_done;       // accumulating products of this block issued and complete
  uint32_t tmem_base;
};

constexpr int kBwdThreads = 384;          // warps 0-3: TMA / MMA / TMEM allocator / spare; warps 4-11: row threads
constexpr int kBwdBlk = 64;               // streamed keys (dQ kernel) / queries (dK,dV kernel) per step
constexpr int kHalfTile = kBwdBlk * 64 * 2;   // 64 x 64 x 2 B = 8 KiB

template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}

// Both kernels: 256 TMEM columns and < 113 KiB of shared memory per CTA so that two CTAs share an SM
// (the second hides the first one's TMA -> MMA -> row-thread -> MMA latency chain), streamed operand
// blocks of 64 rows, and EIGHT row warps: two threads per row, each owning 32 of the 64 streamed
// columns (the backward needs no row reductions — L and D are known — so the split is free), which
// doubles the warps available to hide MUFU / TMEM-load latency.  setmaxnreg moves registers from the
// TMA/MMA warpgroup to the two row warpgroups.

// ============================================================================ dQ
// smem: Q 16K | dO 16K | ring 2 x (K 8K + V 8K) | dS 16K | bars.   TMEM: S [0,64) dP [64,128) dQ [128,192)
constexpr int kDqSmem = 2 * kTile + 2 * 2 * kHalfTile + kTile + 256;

template <bool kBf16, bool kDrop>
__global__ void __launch_bounds__(kBwdThreads, 2)
attention_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                        const __grid_constant__ CUtensorMap tmap_dq, const AttnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t off_q = 0, off_do = kTile, off_ring = 2 * kTile, off_ds = off_ring + 4 * kHalfTile,
                     off_bar = off_ds + kTile;
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + off_bar);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q0 = blockIdx.x * kAttnBQ, head = blockIdx.y, b = blockIdx.z;
  const uint32_t nblk = (a.sk + kBwdBlk - 1) / kBwdBlk;        // 64-key steps
  const uint32_t nlive = (a.sk + kAttnBK - 1) / kAttnBK;       // the live map is per 128 keys
  const int32_t col_h = static_cast<int32_t>(head * kAttnHeadDim);
  const uint8_t* k_live = a.k_live ? a.k_live + static_cast<size_t>(b) * nlive : nullptr;
  auto next_live = [&](uint32_t j) {
    while (j < nblk && k_live && k_live[j >> 1] == 0) ++j;
    return j;
  };
  const bool cta_dead = a.q_live && a.q_live[static_cast<size_t>(b) * gridDim.x + blockIdx.x] == 0;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->fixed_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->ring_full[s]), 1);
      mbar_init(smem_u32(&bars->ring_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->sdp_full), 1);
    mbar_init(smem_u32(&bars->sdp_empty), 8);
    mbar_init(smem_u32(&bars->pds_full), 8);
    mbar_init(smem_u32(&bars->acc_done), 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bars->tmem_base), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_dq = tmem_base + 128;

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      if (lane == 0 && !cta_dead) {
        const uint32_t fb = smem_u32(&bars->fixed_full);
        mbar_arrive_expect_tx(fb, 2 * kTile);
        tma_load_3d(smem_base + off_q, &tmap_q, fb, col_h, static_cast<int32_t>(q0), static_cast<int32_t>(b), kEvictNormal);
        tma_load_3d(smem_base + off_do, &tmap_do, fb, col_h, static_cast<int32_t>(q0), static_cast<int32_t>(b), kEvictNormal);
        uint32_t stage = 0, phase = 0;
        for (uint32_t j = next_live(0); j < nblk; j = next_live(j + 1)) {
          mbar_wait_backoff<100>(smem_u32(&bars->ring_empty[stage]), phase ^ 1);   // sleeps between polls: a tight loop takes issue slots from the row warps
          const uint32_t fbar = smem_u32(&bars->ring_full[stage]);
          mbar_arrive_expect_tx(fbar, 2 * kHalfTile);
          const uint32_t dst = smem_base + off_ring + stage * 2 * kHalfTile;
          tma_load_3d(dst, &tmap_k, fbar, col_h, static_cast<int32_t>(j * kBwdBlk), static_cast<int32_t>(b), kEvictLast);
          tma_load_3d(dst + kHalfTile, &tmap_v, fbar, col_h, static_cast<int32_t>(j * kBwdBlk), static_cast<int32_t>(b), kEvictLast);
          if (++stage == 2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && !cta_dead) {
        mbar_wait(smem_u32(&bars->fixed_full), 0);
        tc_fence_after();
        const uint64_t qdesc = smem_desc_sw128(smem_base + off_q);
        const uint64_t dodesc = smem_desc_sw128(smem_base + off_do);
        uint32_t ld_stage = 0, ld_phase = 0;
        auto issue_sdp = [&](uint32_t n) {   // S = Q.K^T and dP = dO.V^T of the n-th live block
          mbar_wait(smem_u32(&bars->ring_full[ld_stage]), ld_phase);
          mbar_wait(smem_u32(&bars->sdp_empty), (n & 1) ^ 1);
          tc_fence_after();
          const uint32_t kbase = smem_base + off_ring + ld_stage * 2 * kHalfTile;
          const uint64_t kdesc = smem_desc_sw128(kbase), vdesc = smem_desc_sw128(kbase + kHalfTile);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16_ss(tmem_s, qdesc + static_cast<uint64_t>(kk * 2), kdesc + static_cast<uint64_t>(kk * 2),
                       a.idesc_s, kk != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16_ss(tmem_dp, dodesc + static_cast<uint64_t>(kk * 2), vdesc + static_cast<uint64_t>(kk * 2),
                       a.idesc_s, kk != 0 ? 1u : 0u);
          mma_commit(smem_u32(&bars->sdp_full));
          if (++ld_stage == 2) {
            ld_stage = 0;
            ld_phase ^= 1;
          }
        };
        issue_sdp(0);
        uint32_t stage = 0, n = 0;
        for (uint32_t j = next_live(0); j < nblk; ++n) {
          j = next_live(j + 1);
          if (j < nblk) issue_sdp(n + 1);          // overlaps the row threads' work on block n
          mbar_wait_backoff<96>(smem_u32(&bars->pds_full), n & 1);
          tc_fence_after();
          const uint32_t kbase = smem_base + off_ring + stage * 2 * kHalfTile;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {   // dQ += dS . K   (A = dS K-major, B = K MN-major)
            const uint64_t dsdesc = smem_desc_sw128(smem_base + off_ds) + static_cast<uint64_t>(ks * 2);
            const uint64_t kmn = smem_desc_sw128_mn(kbase + ks * 2048, 1024, 1024);
            mma_f16_ss(tmem_dq, dsdesc, kmn, a.idesc_o, (n != 0 || ks != 0) ? 1u : 0u);
          }
          mma_commit(smem_u32(&bars->ring_empty[stage]));
          mma_commit(smem_u32(&bars->acc_done));
          if (++stage == 2) stage = 0;
        }
      }
    }
  } else {
    reg_alloc<96>();    // pool = 384 x 80 registers: 128 x 40 + 256 x 96 fits, 104 would deadlock
    const uint32_t quad = warp & 3, row = quad * 32 + lane, qi = q0 + row;
    const uint32_t ch = (warp - 4) >> 2;            // which 32 of the 64 streamed columns this thread owns
    const uint32_t lane_tmem = (quad * 32) << 16;
    const bool row_active = qi < a.sq;
    uint8_t* ds_row = smem + off_ds + row * 128u;
    if (cta_dead) {
      uint8_t* o_row = smem + off_q + row * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(o_row + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
    } else {
      const bool q_is_pad = row_active && a.q_pad && a.q_pad[static_cast<size_t>(b) * a.sq + qi] != 0;
      const size_t stat = (static_cast<size_t>(b) * a.heads + head) * a.sq + (row_active ? qi : 0);
      const float lse2 = a.lse[stat] * kLog2e;
      const float dsum = a.dvec[stat];
      const bool row_dead = q_is_pad || !row_active;     // no gradient reaches any score of this row
      uint32_t drop_row = 0;
      if constexpr (kDrop)
        drop_row = dropout_row_hash(a.drop.key_a, a.drop.key_b, (static_cast<uint64_t>(b) * a.heads + head) * a.sq + qi);
      // a warp whose 32 query rows are all padding (the tail of the sequence's last tile) contributes dS = 0 to
      // every block: it writes its rows of the dS buffer once and then only keeps the barrier protocol going
      const bool warp_dead = __all_sync(kFull, row_dead);
      if (warp_dead) {
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(ds_row + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
      }
      uint32_t n = 0;
      for (uint32_t j = next_live(0); j < nblk; j = next_live(j + 1), ++n) {
        if (warp_dead) {
          mbar_wait(smem_u32(&bars->sdp_full), n & 1);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(smem_u32(&bars->sdp_empty));
            mbar_arrive(smem_u32(&bars->pds_full));
          }
          // the next score tile can land before the working warps have finished this block: without this wait the
          // warp would run a block ahead and its next arrival would complete THIS block's pds_full phase early
          mbar_wait_backoff<64>(smem_u32(&bars->pds_full), n & 1);
          continue;
        }
        const uint32_t kb0 = j * kBwdBlk;
        uint32_t km[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t idx = kb0 + c * 32 + lane;
          const bool f = a.k_pad && idx < a.sk && a.k_pad[static_cast<size_t>(b) * a.sk + idx] != 0;
          km[c] = __ballot_sync(kFull, f);
        }
        const uint32_t valid = min(static_cast<uint32_t>(kBwdBlk), a.sk - kb0);
        const bool plain = !row_dead && valid == kBwdBlk && (km[0] | km[1]) == 0u &&
                           !(a.causal && kb0 + kBwdBlk - 1 > qi);
        mbar_wait(smem_u32(&bars->sdp_full), n & 1);
        tc_fence_after();
        uint32_t s[32], dp[32];
        tmem_ld_32x32b_x32(tmem_s + lane_tmem + ch * 32, s);
        tmem_ld_32x32b_x32(tmem_dp + lane_tmem + ch * 32, dp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->sdp_empty));   // S/dP may be overwritten now
        if (n > 0) mbar_wait(smem_u32(&bars->acc_done), (n - 1) & 1);   // dS buffer consumed
        const uint32_t kmw = ch ? km[1] : km[0];
        // Two copies of the element loop behind ONE warp-uniform branch: a block without padding, causal cut or
        // ragged end in any of the warp's 32 rows (the common case) runs without the per-element mask logic.
        // dS is written WITHOUT the softmax scale; dQ is multiplied by it once, at the end.
        auto block = [&](auto fast_c) {
          constexpr bool kFast = decltype(fast_c)::value;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float ds[8];
            uint32_t cb[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if constexpr (kDrop) {   // column hashes of this group's 8 keys: the same address in every lane
              const uint4* src = reinterpret_cast<const uint4*>(a.drop.colhash + kb0 + ch * 32 + g * 8);
              const uint4 c0 = __ldg(src), c1 = __ldg(src + 1);
              cb[0] = c0.x; cb[1] = c0.y; cb[2] = c0.z; cb[3] = c0.w;
              cb[4] = c1.x; cb[5] = c1.y; cb[6] = c1.z; cb[7] = c1.w;
            }
            if constexpr (kFast) {
              // the mask-free loop on PAIRS: scale-and-shift, the dropout rescale and the product are one packed
              // instruction per two elements each
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const int c = g * 8 + i;
                float t0, t1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), f2_splat(a.scale_log2),
                                 f2_splat(-lse2)), t0, t1);
                const uint64_t p2 = f2_pack(ex2(t0), ex2(t1));
                uint64_t x2;
                if constexpr (kDrop) {
                  const float m0 = dropout_keep(drop_row, cb[i], a.drop.threshold) ? a.drop.inv_keep : 0.f;
                  const float m1 = dropout_keep(drop_row, cb[i + 1], a.drop.threshold) ? a.drop.inv_keep : 0.f;
                  x2 = f2_fma(f2_pack(__uint_as_float(dp[c]), __uint_as_float(dp[c + 1])), f2_pack(m0, m1),
                              f2_splat(-dsum));
                } else {
                  x2 = f2_add(f2_pack(__uint_as_float(dp[c]), __uint_as_float(dp[c + 1])), f2_splat(-dsum));
                }
                f2_unpack(f2_mul(p2, x2), ds[i], ds[i + 1]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int c = g * 8 + i;
                const float p = ex2(fmaf(__uint_as_float(s[c]), a.scale_log2, -lse2));
                float x;
                if constexpr (kDrop) {   // dP = keep / (1 - p_drop) * (dO . V^T): the forward's mask, regenerated
                  const float m = dropout_keep(drop_row, cb[i], a.drop.threshold) ? a.drop.inv_keep : 0.f;
                  x = fmaf(__uint_as_float(dp[c]), m, -dsum);
                } else {
                  x = __uint_as_float(dp[c]) - dsum;
                }
                float d = p * x;
                if (!plain) {
                  const uint32_t cc = ch * 32 + c;
                  const bool masked = row_dead || ((kmw >> c) & 1u) || (a.causal && kb0 + cc > qi) || cc >= valid;
                  d = masked ? 0.f : d;
                }
                ds[i] = d;
              }
            }
            const uint32_t phys = (static_cast<uint32_t>(ch * 4 + g) ^ (row & 7u)) * 16u;
            *reinterpret_cast<uint4*>(ds_row + phys) =
                make_uint4(pack2<kBf16>(ds[0], ds[1]), pack2<kBf16>(ds[2], ds[3]), pack2<kBf16>(ds[4], ds[5]),
                           pack2<kBf16>(ds[6], ds[7]));
          }
        };
        if (__all_sync(kFull, plain)) block(std::true_type{});
        else block(std::false_type{});
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->pds_full));
      }
      mbar_wait(smem_u32(&bars->acc_done), (n - 1) & 1);
      tc_fence_after();
      uint8_t* o_row = smem + off_q + row * 128u;   // Q is dead now
      {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem_dq + lane_tmem + ch * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * a.scale);   // dQ = scale * dS . K
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t phys = (static_cast<uint32_t>(ch * 4 + g) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(o_row + phys) = make_uint4(
              pack2<kBf16>(__uint_as_float(o[g * 8]), __uint_as_float(o[g * 8 + 1])),
              pack2<kBf16>(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3])),
              pack2<kBf16>(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5])),
              pack2<kBf16>(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7])));
        }
      }
      fence_proxy_async_smem();
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 4 && lane == 0) {
      tma_store_3d(&tmap_dq, smem_base + off_q, col_h, static_cast<int32_t>(q0), static_cast<int32_t>(b));
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// ============================================================================ dK, dV
// smem: K 16K | V 16K | ring 2 x (Q 8K + dO 8K) | P^T 16K | dS^T 16K | stats 8 warps x 384 B | bars
// TMEM: S^T [0,64) dP^T [64,128) dV [128,192) dK [192,256)
constexpr int kDkvSmem = 2 * kTile + 2 * 2 * kHalfTile + 2 * kTile + 3072 + 256;

template <bool kBf16, bool kDrop>
__global__ void __launch_bounds__(kBwdThreads, 2)
attention_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                         const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                         const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
                         const AttnBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t off_k = 0, off_v = kTile, off_ring = 2 * kTile, off_p = off_ring + 4 * kHalfTile,
                     off_ds = off_p + kTile, off_stat = off_ds + kTile, off_bar = off_stat + 3072;
  BwdBars* bars = reinterpret_cast<BwdBars*>(smem + off_bar);
  float* stat_smem = reinterpret_cast<float*>(smem + off_stat);   // per row-warp: 32 x {lse*log2e, D, dropout row hash}
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t k0 = blockIdx.x * kAttnBK, head = blockIdx.y, b = blockIdx.z;
  const uint32_t nqblk = (a.sq + kBwdBlk - 1) / kBwdBlk;       // 64-query steps
  const uint32_t nlive = (a.sq + kAttnBQ - 1) / kAttnBQ;
  const int32_t col_h = static_cast<int32_t>(head * kAttnHeadDim);
  const uint8_t* q_live = a.q_live ? a.q_live + static_cast<size_t>(b) * nlive : nullptr;
  auto next_live = [&](uint32_t i) {
    while (i < nqblk && q_live && q_live[i >> 1] == 0) ++i;
    return i;
  };
  const bool cta_dead = a.k_live && a.k_live[static_cast<size_t>(b) * gridDim.x + blockIdx.x] == 0;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->fixed_full), 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->ring_full[s]), 1);
      mbar_init(smem_u32(&bars->ring_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->sdp_full), 1);
    mbar_init(smem_u32(&bars->sdp_empty), 8);
    mbar_init(smem_u32(&bars->pds_full), 8);
    mbar_init(smem_u32(&bars->acc_done), 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bars->tmem_base), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 64, tmem_dv = tmem_base + 128, tmem_dk = tmem_base + 192;

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      if (lane == 0 && !cta_dead) {
        const uint32_t fb = smem_u32(&bars->fixed_full);
        mbar_arrive_expect_tx(fb, 2 * kTile);
        tma_load_3d(smem_base + off_k, &tmap_k, fb, col_h, static_cast<int32_t>(k0), static_cast<int32_t>(b), kEvictNormal);
        tma_load_3d(smem_base + off_v, &tmap_v, fb, col_h, static_cast<int32_t>(k0), static_cast<int32_t>(b), kEvictNormal);
        uint32_t stage = 0, phase = 0;
        for (uint32_t i = next_live(0); i < nqblk; i = next_live(i + 1)) {
          mbar_wait_backoff<100>(smem_u32(&bars->ring_empty[stage]), phase ^ 1);   // sleeps between polls: a tight loop takes issue slots from the row warps
          const uint32_t fbar = smem_u32(&bars->ring_full[stage]);
          mbar_arrive_expect_tx(fbar, 2 * kHalfTile);
          const uint32_t dst = smem_base + off_ring + stage * 2 * kHalfTile;
          tma_load_3d(dst, &tmap_q, fbar, col_h, static_cast<int32_t>(i * kBwdBlk), static_cast<int32_t>(b), kEvictLast);
          tma_load_3d(dst + kHalfTile, &tmap_do, fbar, col_h, static_cast<int32_t>(i * kBwdBlk), static_cast<int32_t>(b), kEvictLast);
          if (++stage == 2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && !cta_dead) {
        mbar_wait(smem_u32(&bars->fixed_full), 0);
        tc_fence_after();
        const uint64_t kdesc = smem_desc_sw128(smem_base + off_k);
        const uint64_t vdesc = smem_desc_sw128(smem_base + off_v);
        uint32_t ld_stage = 0, ld_phase = 0;
        auto issue_sdp = [&](uint32_t n) {   // S^T = K.Q^T and dP^T = V.dO^T  [keys x queries]
          mbar_wait(smem_u32(&bars->ring_full[ld_stage]), ld_phase);
          mbar_wait(smem_u32(&bars->sdp_empty), (n & 1) ^ 1);
          tc_fence_after();
          const uint32_t qbase = smem_base + off_ring + ld_stage * 2 * kHalfTile;
          const uint64_t qdesc = smem_desc_sw128(qbase), dodesc = smem_desc_sw128(qbase + kHalfTile);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16_ss(tmem_s, kdesc + static_cast<uint64_t>(kk * 2), qdesc + static_cast<uint64_t>(kk * 2),
                       a.idesc_s, kk != 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16_ss(tmem_dp, vdesc + static_cast<uint64_t>(kk * 2), dodesc + static_cast<uint64_t>(kk * 2),
                       a.idesc_s, kk != 0 ? 1u : 0u);
          mma_commit(smem_u32(&bars->sdp_full));
          if (++ld_stage == 2) {
            ld_stage = 0;
            ld_phase ^= 1;
          }
        };
        issue_sdp(0);
        uint32_t stage = 0, n = 0;
        for (uint32_t i = next_live(0); i < nqblk; ++n) {
          i = next_live(i + 1);
          if (i < nqblk) issue_sdp(n + 1);
          mbar_wait_backoff<96>(smem_u32(&bars->pds_full), n & 1);
          tc_fence_after();
          const uint32_t qbase = smem_base + off_ring + stage * 2 * kHalfTile;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t koff = static_cast<uint64_t>(ks * 2);
            const uint64_t pdesc = smem_desc_sw128(smem_base + off_p) + koff;
            const uint64_t dsdesc = smem_desc_sw128(smem_base + off_ds) + koff;
            const uint64_t domn = smem_desc_sw128_mn(qbase + kHalfTile + ks * 2048, 1024, 1024);
            const uint64_t qmn = smem_desc_sw128_mn(qbase + ks * 2048, 1024, 1024);
            const uint32_t acc = (n != 0 || ks != 0) ? 1u : 0u;
            mma_f16_ss(tmem_dv, pdesc, domn, a.idesc_o, acc);    // dV += P^T . dO
            mma_f16_ss(tmem_dk, dsdesc, qmn, a.idesc_o, acc);    // dK += dS^T . Q
          }
          mma_commit(smem_u32(&bars->ring_empty[stage]));
          mma_commit(smem_u32(&bars->acc_done));
          if (++stage == 2) stage = 0;
        }
      }
    }
  } else {
    reg_alloc<96>();    // pool = 384 x 80 registers: 128 x 40 + 256 x 96 fits, 104 would deadlock
    const uint32_t quad = warp & 3, row = quad * 32 + lane, kj = k0 + row;   // row == key
    const uint32_t ch = (warp - 4) >> 2;            // which 32 of the 64 streamed columns this thread owns
    const uint32_t lane_tmem = (quad * 32) << 16;
    const bool row_active = kj < a.sk;
    uint8_t* p_row = smem + off_p + row * 128u;
    uint8_t* ds_row = smem + off_ds + row * 128u;
    if (cta_dead) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        *reinterpret_cast<uint4*>(smem + off_k + row * 128u + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(smem + off_v + row * 128u + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async_smem();
    } else {
      const bool k_is_pad = row_active && a.k_pad && a.k_pad[static_cast<size_t>(b) * a.sk + kj] != 0;
      uint32_t drop_col = 0;   // this thread's key column of the dropout plane (table covers roundup(sk, 128))
      if constexpr (kDrop) drop_col = __ldg(a.drop.colhash + kj);
      // a warp whose 32 keys are all padding / past the end contributes P = dS = 0 to every block: it writes its
      // rows of the two buffers once and then only keeps the barrier protocol going
      const bool warp_dead = __all_sync(kFull, !row_active || k_is_pad);
      if (warp_dead) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<uint4*>(p_row + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(ds_row + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
      }
      // per-query statistics of the block's 32 queries this warp walks: every warp stages ITS OWN copy (three
      // coalesced loads, a __syncwarp) — a CTA-wide barrier per block was the kernel's top stall reason
      float* st = stat_smem + (warp - 4) * 96;   // {lse * log2e, D, dropout row hash} x 32
      uint32_t n = 0;
      for (uint32_t i = next_live(0); i < nqblk; i = next_live(i + 1), ++n) {
        if (warp_dead) {
          mbar_wait(smem_u32(&bars->sdp_full), n & 1);
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(smem_u32(&bars->sdp_empty));
            mbar_arrive(smem_u32(&bars->pds_full));
          }
          // the next score tile can land before the working warps have finished this block: without this wait the
          // warp would run a block ahead and its next arrival would complete THIS block's pds_full phase early
          mbar_wait_backoff<64>(smem_u32(&bars->pds_full), n & 1);
          continue;
        }
        const uint32_t qb0 = i * kBwdBlk;
        const uint32_t qcol = qb0 + ch * 32 + lane;          // the query this lane stages
        {
          const size_t sidx = (static_cast<size_t>(b) * a.heads + head) * a.sq + (qcol < a.sq ? qcol : 0);
          const float lse2 = a.lse[sidx] * kLog2e, dsum = a.dvec[sidx];
          __syncwarp();                                       // the previous block's reads are done
          st[lane] = lse2;
          st[32 + lane] = dsum;
          if constexpr (kDrop)
            st[64 + lane] = __uint_as_float(dropout_row_hash(
                a.drop.key_a, a.drop.key_b, (static_cast<uint64_t>(b) * a.heads + head) * a.sq + qcol));
          __syncwarp();
        }
        const uint32_t qmw = __ballot_sync(
            kFull, a.q_pad && qcol < a.sq && a.q_pad[static_cast<size_t>(b) * a.sq + qcol] != 0);
        const uint32_t valid = min(static_cast<uint32_t>(kBwdBlk), a.sq - qb0);
        const bool plain = row_active && !k_is_pad && valid == kBwdBlk && qmw == 0u && !(a.causal && kj > qb0);
        mbar_wait(smem_u32(&bars->sdp_full), n & 1);
        tc_fence_after();
        uint32_t s[32], dp[32];
        tmem_ld_32x32b_x32(tmem_s + lane_tmem + ch * 32, s);
        tmem_ld_32x32b_x32(tmem_dp + lane_tmem + ch * 32, dp);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->sdp_empty));
        if (n > 0) mbar_wait(smem_u32(&bars->acc_done), (n - 1) & 1);
        // As in the dQ kernel: one warp-uniform branch selects the element loop without mask logic when no row of
        // the warp needs any; dS^T carries no softmax scale (dK is multiplied by it once, at the end).
        auto block = [&](auto fast_c) {
          constexpr bool kFast = decltype(fast_c)::value;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8], ds[8];
            const float* lp = st + g * 8;
            const float4 l0 = *reinterpret_cast<const float4*>(lp);
            const float4 l1 = *reinterpret_cast<const float4*>(lp + 4);
            const float4 d0 = *reinterpret_cast<const float4*>(lp + 32);
            const float4 d1 = *reinterpret_cast<const float4*>(lp + 36);
            const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
            uint32_t rh[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if constexpr (kDrop) {
              const uint4 r0 = *reinterpret_cast<const uint4*>(lp + 64);
              const uint4 r1 = *reinterpret_cast<const uint4*>(lp + 68);
              rh[0] = r0.x; rh[1] = r0.y; rh[2] = r0.z; rh[3] = r0.w;
              rh[4] = r1.x; rh[5] = r1.y; rh[6] = r1.z; rh[7] = r1.w;
            }
            if constexpr (kFast) {
#pragma unroll
              for (int i2 = 0; i2 < 8; i2 += 2) {       // the mask-free loop on pairs (packed fp32 instructions)
                const int c = g * 8 + i2;
                float t0, t1;
                f2_unpack(f2_fma(f2_pack(__uint_as_float(s[c]), __uint_as_float(s[c + 1])), f2_splat(a.scale_log2),
                                 f2_pack(-lv[i2], -lv[i2 + 1])), t0, t1);
                const uint64_t p2 = f2_pack(ex2(t0), ex2(t1));
                uint64_t pd2 = p2, x2;
                const uint64_t dp2 = f2_pack(__uint_as_float(dp[c]), __uint_as_float(dp[c + 1]));
                const uint64_t ndv2 = f2_pack(-dv[i2], -dv[i2 + 1]);
                if constexpr (kDrop) {
                  const float m0 = dropout_keep(rh[i2], drop_col, a.drop.threshold) ? a.drop.inv_keep : 0.f;
                  const float m1 = dropout_keep(rh[i2 + 1], drop_col, a.drop.threshold) ? a.drop.inv_keep : 0.f;
                  const uint64_t m2 = f2_pack(m0, m1);
                  pd2 = f2_mul(p2, m2);
                  x2 = f2_fma(dp2, m2, ndv2);
                } else {
                  x2 = f2_add(dp2, ndv2);
                }
                f2_unpack(pd2, pv[i2], pv[i2 + 1]);
                f2_unpack(f2_mul(p2, x2), ds[i2], ds[i2 + 1]);
              }
            }
#pragma unroll
            for (int i2 = 0; i2 < (kFast ? 0 : 8); ++i2) {
              const int c = g * 8 + i2;                 // column inside this thread's 32
              const uint32_t cc = ch * 32 + c;          // query column inside the 64-query block
              bool masked = false;
              float p;
              if constexpr (kFast) {
                p = ex2(fmaf(__uint_as_float(s[c]), a.scale_log2, -lv[i2]));
              } else {
                float t = __uint_as_float(s[c]) * a.scale_log2;
                if (!plain) {
                  masked = k_is_pad || ((qmw >> c) & 1u) || (a.causal && kj > qb0 + cc);
                  t = masked ? kMaskedLog2 : t;
                }
                p = ex2(t - lv[i2]);
                if (!plain) p = (cc < valid && row_active) ? p : 0.f;
              }
              float pd = p;            // what multiplied V in the forward: the dropped, rescaled probability
              float x;
              if constexpr (kDrop) {
                const float m = dropout_keep(rh[i2], drop_col, a.drop.threshold) ? a.drop.inv_keep : 0.f;
                pd = p * m;
                x = fmaf(__uint_as_float(dp[c]), m, -dv[i2]);
              } else {
                x = __uint_as_float(dp[c]) - dv[i2];
              }
              const float d = p * x;
              pv[i2] = pd;
              ds[i2] = masked ? 0.f : d;
            }
            const uint32_t phys = (static_cast<uint32_t>(ch * 4 + g) ^ (row & 7u)) * 16u;
            *reinterpret_cast<uint4*>(p_row + phys) =
                make_uint4(pack2<kBf16>(pv[0], pv[1]), pack2<kBf16>(pv[2], pv[3]), pack2<kBf16>(pv[4], pv[5]),
                           pack2<kBf16>(pv[6], pv[7]));
            *reinterpret_cast<uint4*>(ds_row + phys) =
                make_uint4(pack2<kBf16>(ds[0], ds[1]), pack2<kBf16>(ds[2], ds[3]), pack2<kBf16>(ds[4], ds[5]),
                           pack2<kBf16>(ds[6], ds[7]));
          }
        };
        if (__all_sync(kFull, plain)) block(std::true_type{});
        else block(std::false_type{});
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->pds_full));
      }
      mbar_wait(smem_u32(&bars->acc_done), (n - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int which = 0; which < 2; ++which) {   // 0: dV -> V's tile, 1: dK -> K's tile (both dead now)
        uint8_t* o_row = smem + (which == 0 ? off_v : off_k) + row * 128u;
        const uint32_t taddr = (which == 0 ? tmem_dv : tmem_dk) + lane_tmem + ch * 32;
        uint32_t o[32];
        tmem_ld_32x32b_x32(taddr, o);
        tmem_ld_wait();
        if (which == 1) {   // dK = scale * dS^T . Q
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * a.scale);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t phys = (static_cast<uint32_t>(ch * 4 + g) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(o_row + phys) = make_uint4(
              pack2<kBf16>(__uint_as_float(o[g * 8]), __uint_as_float(o[g * 8 + 1])),
              pack2<kBf16>(__uint_as_float(o[g * 8 + 2]), __uint_as_float(o[g * 8 + 3])),
              pack2<kBf16>(__uint_as_float(o[g * 8 + 4]), __uint_as_float(o[g * 8 + 5])),
              pack2<kBf16>(__uint_as_float(o[g * 8 + 6]), __uint_as_float(o[g * 8 + 7])));
        }
      }
      fence_proxy_async_smem();
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 4 && lane == 0) {
      tma_store_3d(&tmap_dv, smem_base + off_v, col_h, static_cast<int32_t>(k0), static_cast<int32_t>(b));
      tma_store_3d(&tmap_dk, smem_base + off_k, col_h, static_cast<int32_t>(k0), static_cast<int32_t>(b));
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

}  // namespace

template <bool kBf16, bool kDrop>
cudaError_t bwd_prepare_one() {
  cudaError_t e = cudaFuncSetAttribute(attention_bwd_dq_kernel<kBf16, kDrop>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kDqSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attention_bwd_dkv_kernel<kBf16, kDrop>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kDkvSmem);
  return e;
}

cudaError_t attention_bwd_prepare() {
  cudaError_t e = bwd_prepare_one<true, false>();
  if (e == cudaSuccess) e = bwd_prepare_one<false, false>();
  if (e == cudaSuccess) e = bwd_prepare_one<true, true>();
  if (e == cudaSuccess) e = bwd_prepare_one<false, true>();
  return e;
}

cudaError_t launch_attention_bwd_prep(bool bf16, const void* dout, int64_t lddo, const void* out, int64_t ldo,
                                      float* dvec, int batch, int heads, int sq, cudaStream_t stream) {
  const int64_t total = static_cast<int64_t>(batch) * sq * heads;
  if (total <= 0) return cudaSuccess;
  if (total >= (int64_t{1} << 31)) return cudaErrorInvalidValue;
  const int64_t per_block = 256 / 8 * kPrepUnroll;                      // (token, head) units per block
  const unsigned grid = static_cast<unsigned>((total + per_block - 1) / per_block);
  if (bf16)
    attention_bwd_prep_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(dout), lddo,
                                                              static_cast<const uint16_t*>(out), ldo, dvec, batch,
                                                              heads, sq);
  else
    attention_bwd_prep_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(dout), lddo,
                                                               static_cast<const uint16_t*>(out), ldo, dvec, batch,
                                                               heads, sq);
  return cudaGetLastError();
}

template <bool kBf16, bool kDrop>
void bwd_launch_one(const AttnBwdMaps& m, const AttnBwdArgs& a, dim3 gq, dim3 gk, cudaStream_t stream) {
  attention_bwd_dq_kernel<kBf16, kDrop><<<gq, kBwdThreads, kDqSmem, stream>>>(m.q128, m.k64, m.v64, m.do128,
                                                                             m.dq128, a);
  attention_bwd_dkv_kernel<kBf16, kDrop><<<gk, kBwdThreads, kDkvSmem, stream>>>(m.q64, m.k128, m.v128, m.do64,
                                                                               m.dk128, m.dv128, a);
}

cudaError_t launch_attention_bwd(const AttnBwdMaps& m, const AttnBwdArgs& a, bool bf16, cudaStream_t stream) {
  dim3 gq((a.sq + kAttnBQ - 1) / kAttnBQ, a.heads, a.batch);
  dim3 gk((a.sk + kAttnBK - 1) / kAttnBK, a.heads, a.batch);
  const bool drop = a.drop.threshold != 0;
  if (bf16) {
    if (drop) bwd_launch_one<true, true>(m, a, gq, gk, stream);
    else bwd_launch_one<true, false>(m, a, gq, gk, stream);
  } else {
    if (drop) bwd_launch_one<false, true>(m, a, gq, gk, stream);
    else bwd_launch_one<false, false>(m, a, gq, gk, stream);
  }
  return cudaGetLastError();
}

}  // namespace emdr2
