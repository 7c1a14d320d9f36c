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
__global__ void __launch_bounds__(kAttnFwdThreads, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                     const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v,
                     const __grid_constant__ CUtensorMap tmap_o, const AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SW128 tiles need 1024-B alignment
  constexpr uint32_t off_q = 0;
  constexpr uint32_t off_kv = kAttnTileBytes;
  constexpr uint32_t off_p = off_kv + kAttnStages * 2 * kAttnTileBytes;
  constexpr uint32_t off_xchg = off_p + 2 * kAttnTileBytes;   // [2][128] bf16: row-max exchange
  constexpr uint32_t off_bar = off_xchg + kAttnXchgBytes;
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
    mbar_init(smem_u32(&bars->s_empty), 8);
    mbar_init(smem_u32(&bars->p_full), 8);
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
    // ===================================================== softmax + output
    // Eight warps, two threads per query row: thread (row, ch) owns keys [64 ch, 64 ch + 64) of every
    // 128-key block and output columns [32 ch, 32 ch + 32).  The two partial row maxima meet through
    // shared memory (one named barrier per block); the two partial sums meet once at the end.
    reg_alloc<96>();   // pool = 384 x 80 registers: 128 x 40 + 256 x 96
    const uint32_t quad = warp & 3;
    const uint32_t ch = (warp - 4) >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t qi = q0 + row;
    const bool warp_active = q0 + quad * 32 < a.sq;
    const bool row_active = qi < a.sq;
    const bool q_is_pad = row_active && a.q_pad && a.q_pad[static_cast<size_t>(b) * a.sq + qi] != 0;
    const uint32_t lane_tmem = (quad * 32) << 16;
    uint8_t* p_row = smem + off_p + ch * kAttnTileBytes + row * 128u;   // this thread's 64 keys
    // Row-maximum exchange between the two threads of a row: [2 halves][128 rows] bf16, rounded UP.
    // Both threads apply the same rounding to both values, so they agree on m_new, and m_new >= the
    // true maximum (softmax is shift-invariant; the shift cancels in the final division by l).
    __nv_bfloat16* xchg = reinterpret_cast<__nv_bfloat16*>(smem + off_xchg);
    float* xsum = reinterpret_cast<float*>(smem + off_p);               // row-sum exchange: P is dead by then

    if (cta_dead) {   // all-padding query block: zero rows, no attention work
      uint8_t* o_row = smem + off_q + row * 128u;
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(o_row + (ch * 4 + g) * 16) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
    } else {
      float o_acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
      float m_run = __uint_as_float(0xff800000u);  // -inf
      float l_run = 0.f;                           // partial: this thread's 64 keys of every block
      float alpha_prev = 0.f;

      auto accumulate_o = [&](float alpha) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tmem_o + lane_tmem + ch * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] = fmaf(o_acc[i], alpha, __uint_as_float(o[i]));
      };

      uint32_t n = 0;   // running index of the live key block
      for (uint32_t j = next_live(0); j < nblk; j = next_live(j + 1), ++n) {
        const uint32_t kb0 = j * kAttnBK + ch * 64;   // first key of this thread's half
        mbar_wait(smem_u32(&bars->s_full), n & 1);
        tc_fence_after();
        if (n > 0) {   // O_{n-1} has landed and P's buffer is free
          mbar_wait(smem_u32(&bars->o_full), (n - 1) & 1);
          tc_fence_after();
          if (warp_active) accumulate_o(alpha_prev);
        }
        float alpha = 1.f, m_new = m_run;
        uint32_t km[2] = {0u, 0u};
        uint32_t valid = 64;
        bool plain = true, constant = false, causal_row = false;
        if (warp_active) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t idx = kb0 + c * 32 + lane;
            const bool f = a.k_pad && idx < a.sk && a.k_pad[static_cast<size_t>(b) * a.sk + idx] != 0;
            km[c] = __ballot_sync(kFull, f);
          }
          valid = kb0 >= a.sk ? 0u : min(64u, a.sk - kb0);
          causal_row = a.causal && (kb0 + 63 > qi);
          plain = !q_is_pad && !causal_row && valid == 64 && (km[0] | km[1]) == 0u;
          constant = valid == 64 && (q_is_pad || (km[0] & km[1]) == 0xffffffffu || (a.causal && kb0 > qi));
        }
        const uint32_t s_addr = tmem_base + lane_tmem + ch * 64;
        uint32_t t[32];
        // masked log2-domain score of column c (0..31) of chunk `part` from the raw accumulator bits
        auto score = [&](uint32_t raw, uint32_t kmw, int part, int c) -> float {
          float x = __uint_as_float(raw) * a.scale_log2;
          const uint32_t cc = part * 32 + c;
          const bool masked = q_is_pad || ((kmw >> c) & 1u) || (a.causal && kb0 + cc > qi);
          x = masked ? kMaskedLog2 : x;
          return (cc < valid) ? x : __uint_as_float(0xff800000u);
        };
        // ---- pass 1: maximum of this thread's 64 scores
        float mx = __uint_as_float(0xff800000u);
        if (warp_active) {
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            tmem_ld_32x32b_x32(s_addr + part * 32, t);   // .sync.aligned: all lanes, whatever the row needs
            tmem_ld_wait();
            if (constant) {
              mx = kMaskedLog2;
            } else if (plain) {
              float r = __uint_as_float(t[0]);
#pragma unroll
              for (int c = 1; c < 32; ++c) r = fmaxf(r, __uint_as_float(t[c]));
              mx = fmaxf(mx, r * a.scale_log2);          // scale > 0: max commutes with the scaling
            } else {
              const uint32_t kmw = part ? km[1] : km[0];
#pragma unroll
              for (int c = 0; c < 32; ++c) mx = fmaxf(mx, score(t[c], kmw, part, c));
            }
          }
          const __nv_bfloat16 up = __float2bfloat16_ru(mx);
          xchg[ch * 128 + row] = up;
          mx = __bfloat162float(up);
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");   // both halves of every row have their maximum
        if (warp_active) {
          m_new = fmaxf(m_run, fmaxf(mx, __bfloat162float(xchg[(ch ^ 1u) * 128 + row])));
          alpha = ex2(m_run - m_new);
          // ---- pass 2: probabilities of this thread's 64 keys -> P rows in shared memory
          float sum = 0.f;
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            tmem_ld_32x32b_x32(s_addr + part * 32, t);
            tmem_ld_wait();
            const uint32_t kmw = part ? km[1] : km[0];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c = g * 8 + 2 * i;
                float p0, p1;
                if (constant) {
                  p0 = p1 = ex2(kMaskedLog2 - m_new);
                } else if (plain) {
                  p0 = ex2(fmaf(__uint_as_float(t[c]), a.scale_log2, -m_new));
                  p1 = ex2(fmaf(__uint_as_float(t[c + 1]), a.scale_log2, -m_new));
                } else {
                  p0 = ex2(score(t[c], kmw, part, c) - m_new);
                  p1 = ex2(score(t[c + 1], kmw, part, c + 1) - m_new);
                }
                sum += p0 + p1;
                w[i] = pack2<kBf16>(p0, p1);
              }
              const uint32_t phys = (static_cast<uint32_t>(part * 4 + g) ^ (row & 7u)) * 16u;
              *reinterpret_cast<uint4*>(p_row + phys) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
          l_run = fmaf(l_run, alpha, sum);
          m_run = m_new;
          fence_proxy_async_smem();
        }
        alpha_prev = alpha;
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
        xsum[ch * 128 + row] = l_run;
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (warp_active) {
        const float l_tot = l_run + xsum[(ch ^ 1u) * 128 + row];
        const float inv_l = 1.0f / l_tot;
        uint8_t* o_row = smem + off_q + row * 128u;   // Q is dead: every MMA has completed
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack2<kBf16>(o_acc[g * 8 + 2 * i] * inv_l, o_acc[g * 8 + 2 * i + 1] * inv_l);
          const uint32_t phys = (static_cast<uint32_t>(ch * 4 + g) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(o_row + phys) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        if (a.lse && row_active && ch == 0)
          a.lse[(static_cast<size_t>(b) * a.heads + head) * a.sq + qi] = (m_run + log2f(l_tot)) * kLn2;
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 4 && lane == 0) {
      tma_store_3d(&tmap_o, smem_base + off_q, col_h, static_cast<int32_t>(q0), static_cast<int32_t>(b));
      tma_store_commit();
      tma_store_wait<0>();
    }
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
    attention_fwd_kernel<true><<<grid, kAttnFwdThreads, kAttnSmemBytes, stream>>>(tmap_q, tmap_k, tmap_v,
                                                                              tmap_o, args);
  else
    attention_fwd_kernel<false><<<grid, kAttnFwdThreads, kAttnSmemBytes, stream>>>(tmap_q, tmap_k, tmap_v,
                                                                               tmap_o, args);
}

}  // namespace emdr2
