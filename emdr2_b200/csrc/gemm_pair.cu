// CTA-pair (cta_group::2) variant of the forward GEMM, sm_100a only.  See gemm.cuh for the contract;
// this file covers the K-major, 16-bit-output case (bias / GeLU / residual / pre-activation /
// GeLU-backward epilogues) for large M, which is where the step's time goes.
//
// Why: with one CTA per 128x256 tile every SM stages 48 KiB of operands per 64-wide k block
// (85 flop per byte fetched from L2) and only 4 such stages fit beside the output staging — ncu shows
// the tensor pipe 52-78 % busy with L2->SM traffic at 13-17 TB/s.  Here two CTAs on the two SMs of a
// TPC share one 256x256 tile: each stages its own 128 rows of A and HALF of the B tile (32 KiB per
// k block, 128 flop per byte), five stages deep, and CTA rank 0 issues tcgen05.mma.cta_group::2 with
// M = 256, which reads both halves of B from the two shared memories.  Each CTA's 128 accumulator
// rows live in its own TMEM and are drained by its own epilogue warps, exactly as in gemm.cu.
//
// Pair protocol (barriers shared by the pair live in rank 0, reached through mapa):
//   full[s]        rank 0: one arrive.expect_tx(64 KiB) by rank 0's producer; both producers' TMA loads
//                  (cp.async.bulk.tensor ... cta_group::2) complete their bytes on it
//   empty[s]       one per CTA: tcgen05.commit multicast to both after the MMAs that read stage s
//   tmem_full[b]   one per CTA: commit multicast to both after the tile's last MMA
//   tmem_empty[b]  rank 0, count 16: the eight epilogue warps of BOTH CTAs arrive (remote arrive from
//                  rank 1) once they have read accumulator buffer b
#include "gemm.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gelu.cuh"
#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

constexpr int kPairStages = 5;
constexpr int kPairBM = 2 * kGemmBM;                              // rows of the pair's tile
constexpr int kPairStageA = kGemmBM * kGemmBK * 2;                // 16 KiB: this CTA's 128 rows of A
constexpr int kPairStageB = (kGemmBN / 2) * kGemmBK * 2;          // 16 KiB: this CTA's half of B
constexpr int kPairStageBytes = kPairStageA + kPairStageB;        // 32 KiB
constexpr int kPairResBytes = 2 * kGemmBM * 64 * 2;               // residual boxes, one per column half
constexpr int kPairSmemBytes =
    kPairStages * kPairStageBytes + kGemmOutBytes + kPairResBytes + kGemmBarBytes + 1024;
static_assert(kPairSmemBytes <= 227 * 1024, "shared memory budget");

struct PairBars {
  uint64_t full[kPairStages];
  uint64_t empty[kPairStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[2];
  uint32_t tmem_base;
  uint32_t pad_[3];
  alignas(16) uint16_t bias_stage[2][kGemmBN];
};
static_assert(sizeof(PairBars) <= kGemmBarBytes, "barrier block too large");

template <bool kBf16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (kBf16) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
  } else {
    return __half22float2(*reinterpret_cast<const __half2*>(&v));
  }
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


__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool kBf16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r,
                 const __grid_constant__ CUtensorMap tmap_p, const GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr uint32_t off_out = kPairStages * kPairStageBytes;
  constexpr uint32_t off_res = off_out + kGemmOutBytes;
  constexpr uint32_t off_bar = off_res + kPairResBytes;
  PairBars* bars = reinterpret_cast<PairBars*>(smem + off_bar);
  const uint32_t smem_base = smem_u32(smem);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs)
  const uint32_t pair = blockIdx.x >> 1;
  const uint32_t num_pairs = gridDim.x >> 1;
  const uint32_t pair_tiles_m = (a.M + kPairBM - 1) / kPairBM;
  const uint32_t num_work = pair_tiles_m * a.tiles_n;    // 256 x 256 tiles, N fastest
  const uint32_t num_kb = (a.K + kGemmBK - 1) / kGemmBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->tmem_full[b]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[b]), 16);     // 8 epilogue warps of each CTA (used in rank 0)
      mbar_init(smem_u32(&bars->res_full[b]), 1);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_d);
    if (a.flags & (kGemmResidual | kGemmGeluBwd)) prefetch_tmap(&tmap_r);
  }
  if (warp == 2) {     // same warp in both CTAs: the pair allocation is collective
    tmem_alloc_pair(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();      // both CTAs' barriers are initialised before anything reaches across
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (lane == 0) {
      const uint32_t full0 = mapa_cluster(smem_u32(&bars->full[0]), 0);   // the leader's full[] array
      uint32_t stage = 0, phase = 0;
      for (uint32_t w = pair; w < num_work; w += num_pairs) {
        const int32_t m0 = static_cast<int32_t>((w / a.tiles_n) * kPairBM + rank * kGemmBM);
        const int32_t n0 = static_cast<int32_t>((w % a.tiles_n) * kGemmBN + rank * (kGemmBN / 2));
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(smem_u32(&bars->full[stage]), 2 * kPairStageBytes);
          const uint32_t sa = smem_base + stage * kPairStageBytes;
          const uint32_t fbar = full0 + stage * 8;
          const int32_t k0 = static_cast<int32_t>(kb * kGemmBK);
          tma_load_2d_pair(sa, &tmap_a, fbar, k0, m0, kEvictNormal);
          tma_load_2d_pair(sa + kPairStageA, &tmap_b, fbar, k0, n0, kEvictLast);
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (uint32_t w = pair; w < num_work; w += num_pairs, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(smem_u32(&bars->tmem_empty[buf]), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kGemmBN;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kPairStageBytes;
#pragma unroll
          for (int kk = 0; kk < kGemmBK / 16; ++kk) {
            const uint64_t adesc = smem_desc_sw128(sa) + static_cast<uint64_t>(kk * 2);
            const uint64_t bdesc = smem_desc_sw128(sa + kPairStageA) + static_cast<uint64_t>(kk * 2);
            mma_f16_ss_pair(d_tmem, adesc, bdesc, a.idesc, (kb != 0 || kk != 0) ? 1u : 0u);
          }
          mma_commit_pair(smem_u32(&bars->empty[stage]), 3);     // stage free in both CTAs
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit_pair(smem_u32(&bars->tmem_full[buf]), 3);     // accumulators ready in both CTAs
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue (both CTAs, own 128 rows)
    const uint32_t quad = warp & 3;
    const uint32_t hh = (warp - 4) >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t stage_off = off_out + hh * (kGemmBM * 64 * 2);
    uint8_t* stage_ptr = smem + stage_off;
    // The residual (or GeLU-backward aux) box of a chunk lands in its OWN buffer, so its TMA load for
    // chunk c+1 flies while chunk c is still being stored (in gemm.cu it shares the output staging
    // buffer and each chunk pays store-drain + load latency back to back).
    const uint32_t res_off = off_res + hh * (kGemmBM * 64 * 2);
    const uint8_t* res_ptr = smem + res_off;
    const bool issuer = (warp - 4) % 4 == 0 && lane == 0;
    const uint32_t bar_id = 1 + hh;
    const bool has_bias = (a.flags & kGemmBias) != 0;
    const bool has_gelu = (a.flags & kGemmGelu) != 0;
    const bool has_res = (a.flags & (kGemmResidual | kGemmGeluBwd)) != 0;
    const bool gelu_bwd = (a.flags & kGemmGeluBwd) != 0;
    const bool store_pre = (a.flags & kGemmPreact) != 0;
    const uint16_t* bias = static_cast<const uint16_t*>(a.bias);
    const uint32_t res_bar = smem_u32(&bars->res_full[hh]);
    const uint32_t tmem_empty0 = mapa_cluster(smem_u32(&bars->tmem_empty[0]), 0);
    uint32_t res_phase = 0;

    auto tile_m0 = [&](uint32_t w) { return (w / a.tiles_n) * kPairBM + rank * kGemmBM; };
    auto chunk_live = [&](uint32_t w, uint32_t ch) {
      return w < num_work && (w % a.tiles_n) * kGemmBN + hh * 128 + ch * 64 < a.N;
    };
    auto load_residual = [&](uint32_t w, uint32_t ch) {
      mbar_arrive_expect_tx(res_bar, kGemmBM * 64 * 2);
      tma_load_2d(smem_base + res_off, &tmap_r, res_bar,
                  static_cast<int32_t>((w % a.tiles_n) * kGemmBN + hh * 128 + ch * 64),
                  static_cast<int32_t>(tile_m0(w)), kEvictNormal);
    };
    // First live chunk at or after (w, ch) in this column half's processing order (ragged N: the
    // chunks of the last column tile may be past the end for one half); w >= num_work if none.
    auto next_live_chunk = [&](uint32_t& w, uint32_t& ch) {
      while (w < num_work) {
        if (ch == 2) {
          w += num_pairs;
          ch = 0;
          continue;
        }
        if (chunk_live(w, ch)) return;
        ++ch;
      }
    };
    if (has_res && issuer) {
      uint32_t fw = pair, fch = 0;
      next_live_chunk(fw, fch);
      if (fw < num_work) load_residual(fw, fch);
    }

    uint32_t it = 0;
    for (uint32_t w = pair; w < num_work; w += num_pairs, ++it) {
      const uint32_t buf = it & 1;
      const uint32_t m0 = tile_m0(w);
      const uint32_t n0 = (w % a.tiles_n) * kGemmBN;
      if (has_bias) {
        const uint32_t e = threadIdx.x - 128;
        bars->bias_stage[buf][e] = n0 + e < a.N ? bias[n0 + e] : static_cast<uint16_t>(0);
      }
      mbar_wait(smem_u32(&bars->tmem_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      if (has_bias) named_bar_sync(3, 256);
#pragma unroll 1
      for (uint32_t ch = 0; ch < 2; ++ch) {
        const uint32_t col0 = hh * 128 + ch * 64;
        const bool live = n0 + col0 < a.N;
        uint32_t v[64];
        if (live) {
          const uint32_t t_addr = tmem_base + ((quad * 32) << 16) + buf * kGemmBN + col0;
          tmem_ld_32x32b_x32(t_addr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32b_x32(t_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
        }
        if (ch == 1) {   // accumulator buffer read: tell the leader's MMA thread (remote from rank 1)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tmem_empty0 + buf * 8);
        }
        if (!live) continue;

        const uint32_t gcol = n0 + col0;
        if (store_pre) {
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const bool col_ok = gcol + g * 8 < a.N;
            const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]);
            if (has_bias && col_ok) {
              const uint4 bv = *reinterpret_cast<const uint4*>(&bars->bias_stage[buf][col0 + g * 8]);
              const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack2<kBf16>(bw[j]);
                x[2 * j] += f.x;
                x[2 * j + 1] += f.y;
              }
            }
            *reinterpret_cast<uint4*>(stage_ptr + row * 128u + phys) =
                make_uint4(pack2<kBf16>(x[0], x[1]), pack2<kBf16>(x[2], x[3]), pack2<kBf16>(x[4], x[5]),
                           pack2<kBf16>(x[6], x[7]));
          }
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (issuer) {
            tma_store_2d(&tmap_p, smem_base + stage_off, static_cast<int32_t>(gcol), static_cast<int32_t>(m0));
            tma_store_commit();
          }
        }
        if (has_res) {
          mbar_wait(res_bar, res_phase);   // this chunk's residual box has landed
          res_phase ^= 1;
        }
        // arithmetic first: the previous chunk's TMA store drains the staging buffer underneath it
        uint32_t packed[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const bool col_ok = gcol + g * 8 < a.N;
          const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]);
          if (has_bias && col_ok) {
            const uint4 bv = *reinterpret_cast<const uint4*>(&bars->bias_stage[buf][col0 + g * 8]);
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack2<kBf16>(bw[j]);
              x[2 * j] += f.x;
              x[2 * j + 1] += f.y;
            }
          }
          if (has_gelu) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) gelu_erf_pair(x[j], x[j + 1]);
          }
          if (has_res) {   // out-of-range rows / columns were zero-filled by the TMA load
            const uint4 rv = *reinterpret_cast<const uint4*>(res_ptr + row * 128u + phys);
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack2<kBf16>(rw[j]);
              if (gelu_bwd) {
                gelu_erf_grad_pair(f.x, f.y, x[2 * j], x[2 * j + 1]);
              } else {
                x[2 * j] += f.x;
                x[2 * j + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) packed[g * 4 + j] = pack2<kBf16>(x[2 * j], x[2 * j + 1]);
        }
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);   // staging buffer free; every thread has read the residual box
        if (has_res && issuer) {       // fetch the residual of this half's next live chunk
          uint32_t nt = w, nch = ch + 1;
          next_live_chunk(nt, nch);
          if (nt < num_work) load_residual(nt, nch);
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(stage_ptr + row * 128u + phys) =
              make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer) {
          tma_store_2d(&tmap_d, smem_base + stage_off, static_cast<int32_t>(gcol),
                       static_cast<int32_t>(m0));
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  // Neither CTA may leave while its partner can still read its shared memory (pair MMAs), signal its
  // barriers or use the jointly allocated tensor memory.
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

}  // namespace

cudaError_t gemm_pair_prepare() {
  cudaError_t e = cudaFuncSetAttribute(gemm_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kPairSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemm_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kPairSmemBytes);
}

cudaError_t launch_gemm_pair(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b_half, const CUtensorMap& tmap_d,
                             const CUtensorMap& tmap_r, const CUtensorMap& tmap_p, const GemmArgs& args,
                             bool bf16, int grid, cudaStream_t stream) {
  if (grid < 2 || (grid & 1)) return cudaErrorInvalidValue;     // whole CTA pairs only
  if (bf16)
    gemm_pair_kernel<true><<<grid, kGemmThreads, kPairSmemBytes, stream>>>(tmap_a, tmap_b_half, tmap_d, tmap_r,
                                                                           tmap_p, args);
  else
    gemm_pair_kernel<false><<<grid, kGemmThreads, kPairSmemBytes, stream>>>(tmap_a, tmap_b_half, tmap_d, tmap_r,
                                                                            tmap_p, args);
  return cudaGetLastError();
}

}  // namespace emdr2
