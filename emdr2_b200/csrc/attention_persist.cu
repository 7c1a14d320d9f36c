// Persistent fused attention forward, sm_100a only, head dim 64.  See attention.cuh for the contract
// (mask semantics, padding skip) — this file is the production forward kernel.
//
// Work item = (128-query block, head, batch entry); CTAs are persistent (two per SM) and walk the
// item list with a grid stride, so the TMA producer runs ahead into the next item (Q is
// double-buffered, K/V stream through one ring across item boundaries) and the output store of
// item i overlaps item i+1.  Per 128-key block:
//
//   S = Q·Kᵀ            tcgen05.mma (SS), fp32 in TMEM columns [0,128)
//   softmax             4 warps, one thread per query row, all 128 scores in registers (one
//                       tcgen05.ld pass): mask -> running max -> exp2 -> 16-bit P
//   P                   written back to TMEM columns [0,64) over the scores it came from
//                       (tcgen05.st); never touches shared memory
//   O += P·V            tcgen05.mma with the A operand in TMEM (TS form), V consumed MN-major from
//                       its TMA tile; O accumulates in TMEM columns [128,192) across the key blocks
//
// The running maximum is only raised when a block exceeds it by more than 2^8 (the sums and O stay
// exact in fp32; P stays well inside the 16-bit range), so the O rescale — a TMEM round trip — is
// rare after the first block.  tcgen05.mma executes in issue order: S(n+1) is issued right behind
// P·V(n) although it overwrites P(n), and `s_full(n+1)` therefore also tells the softmax warps that
// P·V(n) has landed in O.
#include "attention.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr float kMaskedLog2 = -10000.0f * 1.4426950408889634f;   // masked_fill value, log2 domain
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegInf = -__builtin_huge_valf();
constexpr float kLazyRescale = 8.0f;    // log2 units the running max may lag behind

struct PBars {
  uint64_t q_full[2];
  uint64_t q_empty[2];
  uint64_t kv_full[kAttnStages];
  uint64_t kv_empty[kAttnStages];
  uint64_t s_full;
  uint64_t p_full;
  uint64_t o_full;
  uint32_t tmem_base;
};
static_assert(sizeof(PBars) <= kAttnBarBytes, "barrier block too large");

constexpr uint32_t kOffQ = 0;                                            // 2 x 16 KiB
constexpr uint32_t kOffKV = 2 * kAttnTileBytes;                          // ring of (K, V) pairs
constexpr uint32_t kOffO = kOffKV + kAttnStages * 2 * kAttnTileBytes;    // output staging
constexpr uint32_t kOffBar = kOffO + kAttnTileBytes;
constexpr int kPersistSmemBytes = kOffBar + kAttnBarBytes;
static_assert(kPersistSmemBytes <= 113 * 1024, "two CTAs per SM must fit");

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

template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}

struct Item {
  uint32_t qb, head, b;
};

template <bool kBf16, bool kDrop>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                const __grid_constant__ CUtensorMap tmap_k,
                                const __grid_constant__ CUtensorMap tmap_v,
                                const __grid_constant__ CUtensorMap tmap_o, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SW128 tiles need 1024-B alignment
  PBars* bars = reinterpret_cast<PBars*>(smem + kOffBar);
  const uint32_t smem_base = smem_u32(smem);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nqb = (a.sq + kAttnBQ - 1) / kAttnBQ;
  const uint32_t nblk = (a.sk + kAttnBK - 1) / kAttnBK;
  const uint32_t items = nqb * a.heads * a.batch;
  auto decode = [&](uint32_t item) {
    Item it;
    it.qb = item % nqb;
    const uint32_t r = item / nqb;
    it.head = r % a.heads;
    it.b = r / a.heads;
    return it;
  };
  // Padding skip (see attention.cuh): dead key blocks are neither loaded nor multiplied; a dead query
  // block is stored as zeros.  All three roles evaluate these identically.
  auto item_dead = [&](const Item& it) {
    return a.q_live && a.q_live[static_cast<size_t>(it.b) * nqb + it.qb] == 0;
  };
  auto next_live = [&](const uint8_t* k_live, uint32_t j) {   // first live key block >= j (nblk if none)
    while (j < nblk && k_live && k_live[j] == 0) ++j;
    return j;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->q_full[s]), 1);
      mbar_init(smem_u32(&bars->q_empty[s]), 1);
    }
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), 1);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->p_full), 4);
    mbar_init(smem_u32(&bars->o_full), 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_k);
    prefetch_tmap(&tmap_v);
    prefetch_tmap(&tmap_o);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bars->tmem_base), kAttnTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_o = tmem_base + kAttnBK;

  if (warp < 4) {
    reg_dealloc<40>();   // warpgroup 0 (TMA, MMA, allocator, spare) hands its registers to the softmax warps
    if (warp == 0 && lane == 0) {
      // ===================================================== TMA producer
      uint32_t stage = 0, phase = 0, live_it = 0;
      for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
        const Item it = decode(item);
        if (item_dead(it)) continue;
        const int32_t col_h = static_cast<int32_t>(it.head * kAttnHeadDim);
        const uint32_t buf = live_it & 1;
        // single-thread waits sleep between polls: a tight try_wait loop is always eligible and takes issue slots from
        // the softmax warp that shares its scheduler (ncu: 44 % of this kernel's issued instructions were such polls)
        mbar_wait_backoff<200>(smem_u32(&bars->q_empty[buf]), ((live_it >> 1) & 1) ^ 1);
        const uint32_t qbar = smem_u32(&bars->q_full[buf]);
        mbar_arrive_expect_tx(qbar, kAttnTileBytes);
        tma_load_3d(smem_base + kOffQ + buf * kAttnTileBytes, &tmap_q, qbar, col_h,
                    static_cast<int32_t>(it.qb * kAttnBQ), static_cast<int32_t>(it.b), kEvictNormal);
        const uint8_t* k_live = a.k_live ? a.k_live + static_cast<size_t>(it.b) * nblk : nullptr;
        for (uint32_t j = next_live(k_live, 0); j < nblk; j = next_live(k_live, j + 1)) {
          mbar_wait_backoff<100>(smem_u32(&bars->kv_empty[stage]), phase ^ 1);
          const uint32_t fbar = smem_u32(&bars->kv_full[stage]);
          mbar_arrive_expect_tx(fbar, 2 * kAttnTileBytes);
          const uint32_t dst = smem_base + kOffKV + stage * 2 * kAttnTileBytes;
          tma_load_3d(dst, &tmap_k, fbar, col_h, static_cast<int32_t>(j * kAttnBK),
                      static_cast<int32_t>(it.b), kEvictLast);
          tma_load_3d(dst + kAttnTileBytes, &tmap_v, fbar, col_h, static_cast<int32_t>(j * kAttnBK),
                      static_cast<int32_t>(it.b), kEvictLast);
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        ++live_it;
      }
    } else if (warp == 1 && lane == 0) {
      // ===================================================== MMA issuer
      uint32_t stage = 0, phase = 0;   // K/V ring position of the next S product
      uint32_t blk = 0;                // running count of key blocks (p_full phase)
      uint32_t live_it = 0;
      for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
        const Item it = decode(item);
        if (item_dead(it)) continue;
        const uint32_t buf = live_it & 1;
        mbar_wait(smem_u32(&bars->q_full[buf]), (live_it >> 1) & 1);
        tc_fence_after();
        const uint64_t qdesc = smem_desc_sw128(smem_base + kOffQ + buf * kAttnTileBytes);
        const uint8_t* k_live = a.k_live ? a.k_live + static_cast<size_t>(it.b) * nblk : nullptr;
        auto issue_s = [&]() {
          mbar_wait(smem_u32(&bars->kv_full[stage]), phase);
          tc_fence_after();
          const uint64_t kdesc = smem_desc_sw128(smem_base + kOffKV + stage * 2 * kAttnTileBytes);
#pragma unroll
          for (int kk = 0; kk < kAttnHeadDim / 16; ++kk)
            mma_f16_ss(tmem_base, qdesc + static_cast<uint64_t>(kk * 2),
                       kdesc + static_cast<uint64_t>(kk * 2), a.idesc_s, kk != 0 ? 1u : 0u);
          mma_commit(smem_u32(&bars->s_full));
        };
        issue_s();
        bool first = true;
        for (uint32_t j = next_live(k_live, 0); j < nblk;) {
          const uint32_t jn = next_live(k_live, j + 1);
          mbar_wait(smem_u32(&bars->p_full), blk & 1);
          tc_fence_after();
          const uint32_t vbase = smem_base + kOffKV + stage * 2 * kAttnTileBytes + kAttnTileBytes;
#pragma unroll
          for (int ks = 0; ks < kAttnBK / 16; ++ks) {
            // A = P in TMEM: row = lane, 16 keys = 8 packed 32-bit columns; B = V MN-major, 16 keys =
            // two 8-row groups of 1024 B
            const uint64_t vdesc = smem_desc_sw128_mn(vbase + ks * 2048, 1024, 1024);
            mma_f16_ts(tmem_o, tmem_base + ks * 8, vdesc, a.idesc_o, (!first || ks != 0) ? 1u : 0u);
          }
          mma_commit(smem_u32(&bars->kv_empty[stage]));
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1;
          }
          ++blk;
          first = false;
          if (jn < nblk) {
            issue_s();
          } else {
            mma_commit(smem_u32(&bars->q_empty[buf]));
            mma_commit(smem_u32(&bars->o_full));
          }
          j = jn;
        }
        ++live_it;
      }
    }
  } else {
    // ===================================================== softmax + output (one thread per row)
    reg_alloc<216>();
    const uint32_t quad = warp & 3;
    const uint32_t row = quad * 32 + lane;
    const uint32_t lane_tmem = (quad * 32) << 16;
    uint8_t* o_row = smem + kOffO + row * 128u;
    const bool store_thread = (warp == 4 && lane == 0);
    uint32_t blk = 0, live_it = 0;
    bool store_pending = false;

    for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
      const Item it = decode(item);
      const uint32_t q0 = it.qb * kAttnBQ;
      const uint32_t qi = q0 + row;
      const int32_t col_h = static_cast<int32_t>(it.head * kAttnHeadDim);
      const bool warp_active = q0 + quad * 32 < a.sq;
      const bool row_active = qi < a.sq;
      uint4 out[8];

      if (item_dead(it)) {
#pragma unroll
        for (int g = 0; g < 8; ++g) out[g] = make_uint4(0u, 0u, 0u, 0u);
      } else {
        const bool q_is_pad = row_active && a.q_pad && a.q_pad[static_cast<size_t>(it.b) * a.sq + qi] != 0;
        const uint8_t* k_live = a.k_live ? a.k_live + static_cast<size_t>(it.b) * nblk : nullptr;
        uint32_t drop_row = 0;
        if constexpr (kDrop)
          drop_row = dropout_row_hash(a.drop.key_a, a.drop.key_b,
                                      (static_cast<uint64_t>(it.b) * a.heads + it.head) * a.sq + qi);
        float m_run = kNegInf;
        float l_run = 0.f;
        bool first = true;
        for (uint32_t j = next_live(k_live, 0); j < nblk; j = next_live(k_live, j + 1), ++blk) {
          const uint32_t kb0 = j * kAttnBK;
          // key-side mask bits of this block, 32 keys per word (issued before the wait: global loads)
          uint32_t km[4] = {0u, 0u, 0u, 0u};
          if (warp_active && a.k_pad) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t idx = kb0 + c * 32 + lane;
              const bool f = idx < a.sk && a.k_pad[static_cast<size_t>(it.b) * a.sk + idx] != 0;
              km[c] = __ballot_sync(kFull, f);
            }
          }
          mbar_wait(smem_u32(&bars->s_full), blk & 1);
          tc_fence_after();
          if (warp_active) {
            const uint32_t valid = min(static_cast<uint32_t>(kAttnBK), a.sk - kb0);
            const bool causal_row = a.causal && (kb0 + kAttnBK - 1 > qi);
            const bool plain = !q_is_pad && !causal_row && valid == kAttnBK && (km[0] | km[1] | km[2] | km[3]) == 0u;
            uint32_t t[kAttnBK];   // scores as raw fp32 bits (tcgen05.ld output registers)
            const uint32_t s_addr = tmem_base + lane_tmem;
            tmem_ld_32x32b_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&t[0]));
            tmem_ld_32x32b_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&t[32]));
            tmem_ld_32x32b_x32(s_addr + 64, *reinterpret_cast<uint32_t(*)[32]>(&t[64]));
            tmem_ld_32x32b_x32(s_addr + 96, *reinterpret_cast<uint32_t(*)[32]>(&t[96]));
            tmem_ld_wait();
            // ---- maximum of the masked log2-domain scores; p = exp2(t * mul + bias - m) per 32-key chunk.
            // Whole block unmasked for every row of the warp (the common case): raw maximum, one scale.
            // Otherwise each 32-key chunk is classified per row — unmasked, all masked (one constant),
            // past the end of the keys (-inf) — and keeps its raw accumulators; only a chunk that a padding
            // boundary, the causal diagonal or the end of the keys actually CROSSES for some row of the
            // warp is rewritten element by element (it used to be the whole 128-key block).
            float mx;
            float mul_c[4], bias_c[4];
            if (__all_sync(kFull, plain)) {
              float r0 = __uint_as_float(t[0]), r1 = __uint_as_float(t[1]);
#pragma unroll
              for (int c = 2; c < kAttnBK; c += 2) {
                r0 = fmaxf(r0, __uint_as_float(t[c]));
                r1 = fmaxf(r1, __uint_as_float(t[c + 1]));
              }
              mx = fmaxf(r0, r1) * a.scale_log2;   // scale > 0: the maximum commutes with the scaling
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                mul_c[c] = a.scale_log2;
                bias_c[c] = 0.f;
              }
            } else {
              mx = kNegInf;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint32_t c0 = kb0 + c * 32;
                const bool past_end = c0 >= a.sk;
                const bool all_masked = q_is_pad || km[c] == 0xffffffffu || (a.causal && c0 > qi);
                const bool none_masked = !q_is_pad && km[c] == 0u && !(a.causal && c0 + 31 > qi);
                const bool uniform = past_end || (c0 + 32 <= a.sk && (all_masked || none_masked));
                float m;
                if (__any_sync(kFull, !uniform)) {
                  m = kNegInf;
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    float x = __uint_as_float(t[c * 32 + i]) * a.scale_log2;
                    const bool masked = q_is_pad || ((km[c] >> i) & 1u) || (a.causal && c0 + i > qi);
                    x = masked ? kMaskedLog2 : x;
                    x = (c0 + i < a.sk) ? x : kNegInf;
                    t[c * 32 + i] = __float_as_uint(x);
                    m = fmaxf(m, x);
                  }
                  mul_c[c] = 1.0f;
                  bias_c[c] = 0.f;
                } else {
                  float r0 = __uint_as_float(t[c * 32]), r1 = __uint_as_float(t[c * 32 + 1]);
#pragma unroll
                  for (int i = 2; i < 32; i += 2) {
                    r0 = fmaxf(r0, __uint_as_float(t[c * 32 + i]));
                    r1 = fmaxf(r1, __uint_as_float(t[c * 32 + i + 1]));
                  }
                  const bool constant = all_masked && !past_end;
                  m = past_end ? kNegInf : (constant ? kMaskedLog2 : fmaxf(r0, r1) * a.scale_log2);
                  mul_c[c] = (past_end || constant) ? 0.0f : a.scale_log2;
                  bias_c[c] = past_end ? kNegInf : (constant ? kMaskedLog2 : 0.0f);
                }
                mx = fmaxf(mx, m);
              }
            }
            // ---- running maximum: raised only when the block exceeds it by more than 2^8
            const bool raise = first || mx > m_run + kLazyRescale;
            float alpha = 1.0f;
            if (raise) {
              alpha = ex2(m_run - mx);   // 0 for the first block (m_run = -inf)
              m_run = mx;
            }
            l_run *= alpha;
            if (!first && __any_sync(kFull, raise)) {
              // O *= alpha in TMEM (P·V of the previous block has completed: s_full is committed
              // behind it).  Warp-uniform branch: tcgen05.ld/st are .sync.aligned.
#pragma unroll
              for (int c = 0; c < kAttnHeadDim / 16; ++c) {
                uint32_t o[16];
                tmem_ld_32x32b_x16(tmem_o + lane_tmem + c * 16, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st_32x32b_x16(tmem_o + lane_tmem + c * 16, o);
              }
            }
            // ---- uniform part: P = exp2(t * mul - m_run) -> 16-bit pairs -> TMEM columns [0, 64)
            float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
            for (int c = 0; c < kAttnBK / 32; ++c) {
              const float mul = mul_c[c];
              const float negm = bias_c[c] - m_run;
              uint32_t w[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float p0 = ex2(fmaf(__uint_as_float(t[c * 32 + 2 * i]), mul, negm));
                float p1 = ex2(fmaf(__uint_as_float(t[c * 32 + 2 * i + 1]), mul, negm));
                sum0 += p0;      // the normaliser is the sum of the UNDROPPED probabilities
                sum1 += p1;
                if constexpr (kDrop) {
                  // column hashes of keys kb0 + c*32 + 2i, +1: the same address in every lane (one broadcast
                  // load per pair; keys past the end carry p = 0 whatever the table holds)
                  const uint2 cb = __ldg(reinterpret_cast<const uint2*>(a.drop.colhash + kb0 + c * 32 + 2 * i));
                  p0 = dropout_keep(drop_row, cb.x, a.drop.threshold) ? p0 : 0.f;
                  p1 = dropout_keep(drop_row, cb.y, a.drop.threshold) ? p1 : 0.f;
                }
                w[i] = pack2<kBf16>(p0, p1);
              }
              tmem_st_32x32b_x16(tmem_base + lane_tmem + c * 16, w);
            }
            l_run += sum0 + sum1;
            tmem_st_wait();
          }
          first = false;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->p_full));
        }

        // ---- item epilogue: O / l -> 16-bit row
        mbar_wait(smem_u32(&bars->o_full), live_it & 1);
        tc_fence_after();
        ++live_it;
        if (warp_active) {
          const float inv_l = kDrop ? a.drop.inv_keep / l_run : 1.0f / l_run;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tmem_o + lane_tmem + half * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                w[i] = pack2<kBf16>(__uint_as_float(o[g * 8 + 2 * i]) * inv_l,
                                    __uint_as_float(o[g * 8 + 2 * i + 1]) * inv_l);
              out[half * 4 + g] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          if (a.lse && row_active)
            a.lse[(static_cast<size_t>(it.b) * a.heads + it.head) * a.sq + qi] = (m_run + log2f(l_run)) * kLn2;
        } else {
#pragma unroll
          for (int g = 0; g < 8; ++g) out[g] = make_uint4(0u, 0u, 0u, 0u);
        }
      }

      // ---- stage the 128 x 64 output tile and hand it to TMA (rows past sq are clipped by the map)
      if (store_thread && store_pending) tma_store_wait_read<0>();   // previous item's tile has left smem
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
        *reinterpret_cast<uint4*>(o_row + phys) = out[g];
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (store_thread) {
        tma_store_3d(&tmap_o, smem_base + kOffO, col_h, static_cast<int32_t>(q0), static_cast<int32_t>(it.b));
        tma_store_commit();
        store_pending = true;
      }
    }
    if (store_thread && store_pending) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kAttnTmemCols);
}

}  // namespace

cudaError_t attention_persistent_prepare() {
  cudaError_t e = cudaFuncSetAttribute(attention_fwd_persistent_kernel<true, false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attention_fwd_persistent_kernel<false, false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attention_fwd_persistent_kernel<true, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmemBytes);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attention_fwd_persistent_kernel<false, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, kPersistSmemBytes);
  return e;
}

void launch_attention_fwd_persistent(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k,
                                     const CUtensorMap& tmap_v, const CUtensorMap& tmap_o,
                                     const AttnArgs& args, bool bf16, int sm_count, cudaStream_t stream) {
  const uint32_t nqb = (args.sq + kAttnBQ - 1) / kAttnBQ;
  const uint64_t items = static_cast<uint64_t>(nqb) * args.heads * args.batch;
  const uint64_t slots = 2ull * static_cast<uint64_t>(sm_count);
  // Items are numbered query-block fastest and walked with a grid stride.  With padding the query
  // blocks of a sequence differ in cost (trailing blocks are dead or see fewer live key blocks), so a
  // stride that is a multiple of nqb would pin every CTA to one query-block index for the whole launch
  // (296 CTAs, nqb = 4: half of them idle on half-padded batches).  A stride coprime with nqb makes
  // each CTA rotate through all query-block indices while concurrently running CTAs still cover
  // neighbouring items (K/V of one sequence stay hot in L2).
  uint32_t ctas = static_cast<uint32_t>(items < slots ? items : slots);
  if (items > slots) {
    auto gcd = [](uint32_t x, uint32_t y) {
      while (y) {
        const uint32_t t = x % y;
        x = y;
        y = t;
      }
      return x;
    };
    while (ctas > 1 && gcd(ctas, nqb) != 1) --ctas;
  }
  const dim3 grid(ctas);
#define EMDR2_ATTN_LAUNCH(BF, DROP)                                                                     \
  attention_fwd_persistent_kernel<BF, DROP><<<grid, kAttnThreads, kPersistSmemBytes, stream>>>(tmap_q, tmap_k, \
                                                                                                tmap_v, tmap_o, args)
  const bool drop = args.drop.threshold != 0;
  if (bf16) {
    if (drop) EMDR2_ATTN_LAUNCH(true, true);
    else EMDR2_ATTN_LAUNCH(true, false);
  } else {
    if (drop) EMDR2_ATTN_LAUNCH(false, true);
    else EMDR2_ATTN_LAUNCH(false, false);
  }
#undef EMDR2_ATTN_LAUNCH
}

}  // namespace emdr2
