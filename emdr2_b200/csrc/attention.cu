// See attention.cuh for the design. sm_100a only.
#include "attention.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr float kMaskedLog2 = -10000.0f * 1.4426950408889634f;   // masked_fill value, log2 domain
constexpr float kLn2 = 0.6931471805599453f;

struct AttnBars {
  uint64_t q_full;
  uint64_t kv_full[kAttnStages];
  uint64_t kv_empty[kAttnStages];
  uint64_t s_full;
  uint64_t s_empty;
  uint64_t p_full;
  uint64_t o_full;
  uint32_t tmem_base;
};

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

static_assert(sizeof(AttnBars) <= kAttnBarBytes, "barrier block too large");

template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}

template <bool kBf16>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                     const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v,
                     const __grid_constant__ CUtensorMap tmap_o, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SW128 tiles need 1024-B alignment
  constexpr uint32_t off_q = 0;
  constexpr uint32_t off_kv = kAttnTileBytes;
  constexpr uint32_t off_p = off_kv + kAttnStages * 2 * kAttnTileBytes;
  constexpr uint32_t off_bar = off_p + 2 * kAttnTileBytes;
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + off_bar);
  const uint32_t smem_base = smem_u32(smem);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t q0 = blockIdx.x * kAttnBQ;
  const uint32_t head = blockIdx.y;
  const uint32_t b = blockIdx.z;
  const uint32_t nblk = (a.sk + kAttnBK - 1) / kAttnBK;
  // Optional padding skip (results at padding query positions are then zero / partial instead of
  // the reference's uniform average; no non-padding position can observe the difference):
  //   k_live[b, j] == 0: key block j of batch b is all padding -> not loaded, not multiplied (its
  //                      probabilities are exactly 0 for every non-padding query);
  //   q_live[b, i] == 0: query block i is all padding -> the CTA just stores zeros.
  const uint8_t* k_live = a.k_live ? a.k_live + static_cast<size_t>(b) * nblk : nullptr;
  auto next_live = [&](uint32_t j) {   // first live key block >= j (nblk if none)
    while (j < nblk && k_live && k_live[j] == 0) ++j;
    return j;
  };
  const bool cta_dead = a.q_live && a.q_live[static_cast<size_t>(b) * gridDim.x + blockIdx.x] == 0;
  const int32_t col_h = static_cast<int32_t>(head * kAttnHeadDim);

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->q_full), 1);
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), 1);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    mbar_init(smem_u32(&bars->s_full), 1);
    mbar_init(smem_u32(&bars->s_empty), 4);
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
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0 && !cta_dead) {
      const uint32_t qbar = smem_u32(&bars->q_full);
      mbar_arrive_expect_tx(qbar, kAttnTileBytes);
      tma_load_3d(smem_base + off_q, &tmap_q, qbar, col_h, static_cast<int32_t>(q0),
                  static_cast<int32_t>(b), kEvictNormal);
      uint32_t stage = 0, phase = 0;
      for (uint32_t j = next_live(0); j < nblk; j = next_live(j + 1)) {
        mbar_wait(smem_u32(&bars->kv_empty[stage]), phase ^ 1);
        const uint32_t fbar = smem_u32(&bars->kv_full[stage]);
        mbar_arrive_expect_tx(fbar, 2 * kAttnTileBytes);
        const uint32_t dst = smem_base + off_kv + stage * 2 * kAttnTileBytes;
        tma_load_3d(dst, &tmap_k, fbar, col_h, static_cast<int32_t>(j * kAttnBK),
                    static_cast<int32_t>(b), kEvictLast);
        tma_load_3d(dst + kAttnTileBytes, &tmap_v, fbar, col_h, static_cast<int32_t>(j * kAttnBK),
                    static_cast<int32_t>(b), kEvictLast);
        if (++stage == kAttnStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && !cta_dead) {
      mbar_wait(smem_u32(&bars->q_full), 0);
      tc_fence_after();
      const uint64_t qdesc = smem_desc_sw128(smem_base + off_q);
      uint32_t ld_stage = 0, ld_phase = 0;  // stage/phase of the next S block to issue
      auto issue_s = [&](uint32_t n) {   // n = running index of the live block
        mbar_wait(smem_u32(&bars->kv_full[ld_stage]), ld_phase);
        mbar_wait(smem_u32(&bars->s_empty), (n & 1) ^ 1);
        tc_fence_after();
        const uint64_t kdesc = smem_desc_sw128(smem_base + off_kv + ld_stage * 2 * kAttnTileBytes);
#pragma unroll
        for (int kk = 0; kk < kAttnHeadDim / 16; ++kk)
          mma_f16_ss(tmem_base, qdesc + static_cast<uint64_t>(kk * 2),
                     kdesc + static_cast<uint64_t>(kk * 2), a.idesc_s, kk != 0 ? 1u : 0u);
        mma_commit(smem_u32(&bars->s_full));
        if (++ld_stage == kAttnStages) {
          ld_stage = 0;
          ld_phase ^= 1;
        }
      };
      issue_s(0);
      uint32_t stage = 0, n = 0;
      for (uint32_t j = next_live(0); j < nblk; ++n) {
        j = next_live(j + 1);
        if (j < nblk) issue_s(n + 1);
        mbar_wait(smem_u32(&bars->p_full), n & 1);
        tc_fence_after();
        const uint32_t vbase = smem_base + off_kv + stage * 2 * kAttnTileBytes + kAttnTileBytes;
#pragma unroll
        for (int ks = 0; ks < kAttnBK / 16; ++ks) {
          // A = P: K-major, 64-key blocks of 16 KiB, 32 B per 16-key step inside a block
          const uint64_t pdesc = smem_desc_sw128(smem_base + off_p + (ks >> 2) * kAttnTileBytes) +
                                 static_cast<uint64_t>((ks & 3) * 2);
          // B = V: MN-major, 16 keys = two 8-row groups of 1024 B
          const uint64_t vdesc = smem_desc_sw128_mn(vbase + ks * 2048, 1024, 1024);
          mma_f16_ss(tmem_o, pdesc, vdesc, a.idesc_o, ks != 0 ? 1u : 0u);
        }
        mma_commit(smem_u32(&bars->kv_empty[stage]));
        mma_commit(smem_u32(&bars->o_full));
        if (++stage == kAttnStages) stage = 0;
      }
    }
  }
  } else {
    // ===================================================== softmax + output (one thread per row)
    reg_alloc<216>();
    const uint32_t quad = warp & 3;
    const uint32_t row = quad * 32 + lane;
    const uint32_t qi = q0 + row;
    const bool warp_active = q0 + quad * 32 < a.sq;
    const bool row_active = qi < a.sq;
    const bool q_is_pad = row_active && a.q_pad && a.q_pad[static_cast<size_t>(b) * a.sq + qi] != 0;
    const uint32_t lane_tmem = (quad * 32) << 16;
    uint8_t* p_row = smem + off_p + row * 128u;

    float o_acc[kAttnHeadDim];
#pragma unroll
    for (int i = 0; i < kAttnHeadDim; ++i) o_acc[i] = 0.f;
    float m_run = __uint_as_float(0xff800000u);  // -inf
    float l_run = 0.f;
    float alpha_prev = 0.f;

    auto accumulate_o = [&](float alpha) {
      uint32_t o[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld_32x32b_x32(tmem_o + lane_tmem + half * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          o_acc[half * 32 + i] = fmaf(o_acc[half * 32 + i], alpha, __uint_as_float(o[i]));
      }
    };

    if (cta_dead) {   // all-padding query block: zero rows, no attention work
      uint8_t* o_row = smem + off_q + row * 128u;
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<uint4*>(o_row + g * 16) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 4 && lane == 0) {
        tma_store_3d(&tmap_o, smem_base + off_q, col_h, static_cast<int32_t>(q0), static_cast<int32_t>(b));
        tma_store_commit();
        tma_store_wait<0>();
      }
    } else {
    uint32_t n = 0;   // running index of the live key block
    for (uint32_t j = next_live(0); j < nblk; j = next_live(j + 1), ++n) {
      const uint32_t kb0 = j * kAttnBK;
      mbar_wait(smem_u32(&bars->s_full), n & 1);
      tc_fence_after();
      float alpha = 1.f;
      float m_new = m_run;
      uint32_t t[64];   // scores as raw fp32 bits (tcgen05.ld output registers)
      // ---- key-side mask bits of this block (same in every warp), 32 keys per word
      uint32_t km[4] = {0u, 0u, 0u, 0u};
      uint32_t valid = kAttnBK;
      bool plain = true;        // no masking at all in this block for this row
      bool constant = false;    // every score of this row in this block is the masked value
      bool causal_row = false;  // the causal diagonal crosses this block for this row
      if (warp_active) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t idx = kb0 + c * 32 + lane;
          const bool f = a.k_pad && idx < a.sk && a.k_pad[static_cast<size_t>(b) * a.sk + idx] != 0;
          km[c] = __ballot_sync(kFull, f);
        }
        valid = min(static_cast<uint32_t>(kAttnBK), a.sk - kb0);
        causal_row = a.causal && (kb0 + kAttnBK - 1 > qi);
        plain = !q_is_pad && !causal_row && valid == kAttnBK && (km[0] | km[1] | km[2] | km[3]) == 0u;
        constant = valid == kAttnBK && (q_is_pad || (km[0] & km[1] & km[2] & km[3]) == 0xffffffffu ||
                                        (a.causal && kb0 > qi));
      }
      // Loads the scores of keys [half*64, half*64+64) of the block into t as masked log2-domain
      // values and returns their maximum.  Half 0 is read twice (max pass, then exp pass) so that
      // only 64 scores are ever live next to the 64 output accumulators.
      auto load_scores = [&](auto half_c) -> float {
        constexpr int half = decltype(half_c)::value;
        // tcgen05.ld is .sync.aligned: every lane of the warp executes it, whatever its row needs
        const uint32_t s_addr = tmem_base + lane_tmem + half * 64;
        tmem_ld_32x32b_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&t[0]));
        tmem_ld_32x32b_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&t[32]));
        tmem_ld_wait();
        if (constant) {   // per-row (divergent) from here on: nothing to scale, nothing to compare
#pragma unroll
          for (int c = 0; c < 64; ++c) t[c] = __float_as_uint(kMaskedLog2);
          return kMaskedLog2;
        }
        float mx = __uint_as_float(0xff800000u);
        if (plain) {
          // unmasked rows keep the RAW accumulators: scale > 0, so the maximum commutes with the
          // scaling, and write_probs folds the scaling into one FFMA per element
          float r = __uint_as_float(t[0]);
#pragma unroll
          for (int c = 1; c < 64; ++c) r = fmaxf(r, __uint_as_float(t[c]));
          mx = r * a.scale_log2;
        } else if (!causal_row && valid == kAttnBK) {   // key-padding bits only
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const uint32_t cc = half * 64 + c;
            float x = __uint_as_float(t[c]) * a.scale_log2;
            x = (km[cc >> 5] & (1u << (cc & 31))) ? kMaskedLog2 : x;
            t[c] = __float_as_uint(x);
            mx = fmaxf(mx, x);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            const uint32_t cc = half * 64 + c;
            float x = __uint_as_float(t[c]) * a.scale_log2;
            const bool masked = q_is_pad || ((km[cc >> 5] >> (cc & 31)) & 1u) ||
                                (a.causal && kb0 + cc > qi);
            x = masked ? kMaskedLog2 : x;
            x = (cc < valid) ? x : __uint_as_float(0xff800000u);
            t[c] = __float_as_uint(x);
            mx = fmaxf(mx, x);
          }
        }
        return mx;
      };
      // exp2(t - m_new) of the 64 live scores -> 16-bit -> P rows in shared memory; returns the sum
      auto write_probs = [&](auto half_c) -> float {
        constexpr int half = decltype(half_c)::value;
        if (constant) {   // one probability value for the whole row of this block
          const float p = ex2(kMaskedLog2 - m_new);
          const uint32_t w = pack2<kBf16>(p, p);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(p_row + half * kAttnTileBytes + g * 16) = make_uint4(w, w, w, w);
          return 64.f * p;
        }
        float sum = 0.f;
        const float mul = plain ? a.scale_log2 : 1.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // branch-free: unmasked rows hold raw accumulators (mul = scale), masked rows hold
            // already scaled / replaced values (mul = 1)
            const float p0 = ex2(fmaf(__uint_as_float(t[g * 8 + 2 * i]), mul, -m_new));
            const float p1 = ex2(fmaf(__uint_as_float(t[g * 8 + 2 * i + 1]), mul, -m_new));
            sum += p0 + p1;
            w[i] = pack2<kBf16>(p0, p1);
          }
          const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(p_row + half * kAttnTileBytes + phys) =
              make_uint4(w[0], w[1], w[2], w[3]);
        }
        return sum;
      };

      // O_{j-1} must have landed (and P's buffer be free) before P_j is written; folding it in
      // first keeps the score registers and the tcgen05.ld staging registers from overlapping
      if (n > 0) {
        mbar_wait(smem_u32(&bars->o_full), (n - 1) & 1);
        tc_fence_after();
        if (warp_active) accumulate_o(alpha_prev);
      }
      if (warp_active) {
        const float mx0 = load_scores(std::integral_constant<int, 0>{});
        const float mx1 = load_scores(std::integral_constant<int, 1>{});   // t holds the second half
        m_new = fmaxf(m_run, fmaxf(mx0, mx1));
        alpha = ex2(m_run - m_new);
      }
      alpha_prev = alpha;
      if (warp_active) {
        float sum = write_probs(std::integral_constant<int, 1>{});
        load_scores(std::integral_constant<int, 0>{});
        sum += write_probs(std::integral_constant<int, 0>{});
        l_run = fmaf(l_run, alpha, sum);
        m_run = m_new;
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bars->s_empty));
        mbar_arrive(smem_u32(&bars->p_full));
      }
    }

    mbar_wait(smem_u32(&bars->o_full), (n - 1) & 1);   // n >= 1: the host marks at least one block live
    tc_fence_after();
    if (warp_active) {
      accumulate_o(alpha_prev);
      const float inv_l = 1.0f / l_run;
      uint8_t* o_row = smem + off_q + row * 128u;   // Q is dead: every MMA has completed
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pack2<kBf16>(o_acc[g * 8 + 2 * i] * inv_l, o_acc[g * 8 + 2 * i + 1] * inv_l);
        const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
        *reinterpret_cast<uint4*>(o_row + phys) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      fence_proxy_async_smem();
      if (a.lse && row_active)
        a.lse[(static_cast<size_t>(b) * a.heads + head) * a.sq + qi] = (m_run + log2f(l_run)) * kLn2;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 4 && lane == 0) {
      tma_store_3d(&tmap_o, smem_base + off_q, col_h, static_cast<int32_t>(q0),
                   static_cast<int32_t>(b));
      tma_store_commit();
      tma_store_wait<0>();
    }
    }   // !cta_dead
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kAttnTmemCols);
}

}  // namespace

cudaError_t attention_prepare() {
  cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel<true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention_fwd_kernel<false>,
                              cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes);
}

void launch_attention_fwd(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k,
                          const CUtensorMap& tmap_v, const CUtensorMap& tmap_o,
                          const AttnArgs& args, bool bf16, cudaStream_t stream) {
  dim3 grid((args.sq + kAttnBQ - 1) / kAttnBQ, args.heads, args.batch);
  if (bf16)
    attention_fwd_kernel<true><<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tmap_q, tmap_k, tmap_v,
                                                                              tmap_o, args);
  else
    attention_fwd_kernel<false><<<grid, kAttnThreads, kAttnSmemBytes, stream>>>(tmap_q, tmap_k, tmap_v,
                                                                               tmap_o, args);
}

}  // namespace emdr2
