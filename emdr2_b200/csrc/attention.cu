// See attention.cuh for the design. sm_100a only.
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

struct AttnBars {
  uint64_t q_full;
  uint64_t kv_full[kAttnStages];
  uint64_t kv_empty[kAttnStages];
  uint64_t s_full[2];
  uint64_t s_empty[2];
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

template <bool kBf16>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                     const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v,
                     const __grid_constant__ CUtensorMap tmap_o, const AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
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
  const int32_t col_h = static_cast<int32_t>(head * kAttnHeadDim);

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars->q_full), 1);
    for (int s = 0; s < kAttnStages; ++s) {
      mbar_init(smem_u32(&bars->kv_full[s]), 1);
      mbar_init(smem_u32(&bars->kv_empty[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bars->s_full[i]), 1);
      mbar_init(smem_u32(&bars->s_empty[i]), 4);
    }
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
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_o = tmem_base + 2 * kAttnBK;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      const uint32_t qbar = smem_u32(&bars->q_full);
      mbar_arrive_expect_tx(qbar, kAttnTileBytes);
      tma_load_3d(smem_base + off_q, &tmap_q, qbar, col_h, static_cast<int32_t>(q0),
                  static_cast<int32_t>(b), kEvictNormal);
      uint32_t stage = 0, phase = 0;
      for (uint32_t j = 0; j < nblk; ++j) {
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
    if (lane == 0) {
      mbar_wait(smem_u32(&bars->q_full), 0);
      tc_fence_after();
      const uint64_t qdesc = smem_desc_sw128(smem_base + off_q);
      uint32_t ld_stage = 0, ld_phase = 0;  // stage/phase of the next S block to issue
      auto issue_s = [&](uint32_t j) {
        const uint32_t sb = j & 1;
        mbar_wait(smem_u32(&bars->kv_full[ld_stage]), ld_phase);
        mbar_wait(smem_u32(&bars->s_empty[sb]), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t kdesc = smem_desc_sw128(smem_base + off_kv + ld_stage * 2 * kAttnTileBytes);
#pragma unroll
        for (int kk = 0; kk < kAttnHeadDim / 16; ++kk)
          mma_f16_ss(tmem_base + sb * kAttnBK, qdesc + static_cast<uint64_t>(kk * 2),
                     kdesc + static_cast<uint64_t>(kk * 2), a.idesc_s, kk != 0 ? 1u : 0u);
        mma_commit(smem_u32(&bars->s_full[sb]));
        if (++ld_stage == kAttnStages) {
          ld_stage = 0;
          ld_phase ^= 1;
        }
      };
      issue_s(0);
      uint32_t stage = 0;
      for (uint32_t j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        mbar_wait(smem_u32(&bars->p_full), j & 1);
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
  } else if (warp >= 4) {
    // ===================================================== softmax + output (one thread per row)
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

    for (uint32_t j = 0; j < nblk; ++j) {
      const uint32_t sb = j & 1;
      const uint32_t kb0 = j * kAttnBK;
      mbar_wait(smem_u32(&bars->s_full[sb]), (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[kAttnBK];
      float alpha = 1.f;
      if (warp_active) {
        const uint32_t s_addr = tmem_base + lane_tmem + sb * kAttnBK;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          tmem_ld_32x32b_x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->s_empty[sb]));

      if (warp_active) {
        // ---- key-side mask bits of this block (same in every warp), 32 keys per word
        uint32_t km[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t idx = kb0 + c * 32 + lane;
          const bool f = a.k_pad && idx < a.sk && a.k_pad[static_cast<size_t>(b) * a.sk + idx] != 0;
          km[c] = __ballot_sync(kFull, f);
        }
        const uint32_t valid = min(static_cast<uint32_t>(kAttnBK), a.sk - kb0);
        const bool causal_hit = a.causal && (kb0 + kAttnBK - 1 > qi);
        const bool plain = !q_is_pad && !causal_hit && valid == kAttnBK &&
                           (km[0] | km[1] | km[2] | km[3]) == 0u;
        float mx = __uint_as_float(0xff800000u);
        if (plain) {
#pragma unroll
          for (int c = 0; c < kAttnBK; ++c) {
            const float t = __uint_as_float(v[c]) * a.scale_log2;
            v[c] = __float_as_uint(t);
            mx = fmaxf(mx, t);
          }
        } else {
#pragma unroll
          for (int c = 0; c < kAttnBK; ++c) {
            float t = __uint_as_float(v[c]) * a.scale_log2;
            const bool masked = q_is_pad || ((km[c >> 5] >> (c & 31)) & 1u) ||
                                (a.causal && kb0 + c > qi);
            t = masked ? kMaskedLog2 : t;
            t = (static_cast<uint32_t>(c) < valid) ? t : __uint_as_float(0xff800000u);
            v[c] = __float_as_uint(t);
            mx = fmaxf(mx, t);
          }
        }
        const float m_new = fmaxf(m_run, mx);
        alpha = ex2(m_run - m_new);
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kAttnBK; c += 2) {
          const float p0 = ex2(__uint_as_float(v[c]) - m_new);
          const float p1 = ex2(__uint_as_float(v[c + 1]) - m_new);
          sum += p0 + p1;
          v[c >> 1] = pack2<kBf16>(p0, p1);
        }
        l_run = fmaf(l_run, alpha, sum);
        m_run = m_new;
      }

      // O_{j-1} must have landed (and P's buffer be free) before P_j is written
      if (j > 0) {
        mbar_wait(smem_u32(&bars->o_full), (j - 1) & 1);
        tc_fence_after();
        if (warp_active) accumulate_o(alpha_prev);
      }
      alpha_prev = alpha;
      if (warp_active) {
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const uint32_t phys = (static_cast<uint32_t>(g & 7) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(p_row + (g >> 3) * kAttnTileBytes + phys) =
              make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        }
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->p_full));
    }

    mbar_wait(smem_u32(&bars->o_full), (nblk - 1) & 1);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
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
