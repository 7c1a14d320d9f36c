// See gemm.cuh for the design. sm_100a only.
#include "gemm.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gelu.cuh"
#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

struct GemmBars {
  uint64_t full[kGemmStages];
  uint64_t empty[kGemmStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[4];    // aux box landed in the staging buffer of column half 0 / 1 (kEpi 2: quarter 0..3)
  uint32_t tmem_base;
  uint32_t pad_[3];
  // Bias slice of the current tile, double-buffered by tile parity: fetched by the epilogue threads
  // BEFORE they wait for the accumulator, so its global-load latency hides behind that wait (eight
  // dependent __ldg per chunk used to sit on the epilogue's critical path).
  alignas(16) uint16_t bias_stage[2][kGemmBN];
};
static_assert(sizeof(GemmBars) <= kGemmBarBytes, "barrier block too large");

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


template <int kRegs>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs));
}
template <int kRegs>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs));
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// kAMN / kBMN: that operand is stored with the contraction index as the ROW index ([k, m] resp.
// [k, n] row-major, "MN-major"): its tile is fetched as 64x64-element boxes (one per 64-wide MN atom,
// 8 KiB apart) and consumed through an MN-major UMMA descriptor — this is how the backward products
// dX = dY.W and dW = dY^T.X read the forward's tensors without any transpose in memory.
//
// kWide: SIXTEEN epilogue warps (4 TMEM lane quadrants x 4 column quarters of the tile) instead of eight, for the
// epilogues that only write (bias / GeLU / pre-activation; no aux tile, no fp32 accumulation).  The GeLU epilogue is
// a dependent rcp -> polynomial -> ex2 chain of ~12 instructions per element: with two warps per scheduler it ran at
// IPC 0.4 and took longer than the tile's MMAs at K = 768 (ncu: tensor pipe 58 %, the MMA thread polling
// tmem_empty); four warps per scheduler hide that latency.  Each quarter stages 32 columns at a time through an
// 8 KiB box (128 rows x 64 B, 64-byte swizzle; tmap_d / tmap_p carry 32-column boxes), so the four boxes fit in the
// space of the narrow variant's two and the operand ring keeps its four stages.
//
// kEpi: 0 = eight epilogue warps (everything, incl. fp32 accumulation); 1 = sixteen, write-only epilogues; 2 = sixteen
// with an aux tile (residual / GeLU backward): THREE operand stages and four 16 KiB staging boxes (one 128 x 64
// chunk per quarter per tile), the aux box of the quarter's NEXT tile fetched by TMA into the box as soon as this
// tile's store has drained it — a whole tile ahead of its use, where the eight-warp variant serialises store drain
// -> aux load -> use for every chunk (dU = (dA . W2) * GeLU'(u) ran at 0.6 PFLOP/s, epilogue-bound).
template <bool kBf16, bool kAMN, bool kBMN, int kEpi>
__global__ void __launch_bounds__(kEpi != 0 ? kGemmWideThreads : kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_r,
            const __grid_constant__ CUtensorMap tmap_p, const GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  constexpr bool kWide = kEpi != 0;
  constexpr int kStages = kEpi == 2 ? kGemmAuxStages : kGemmStages;
  constexpr uint32_t off_out = kStages * kGemmStageBytes;
  constexpr uint32_t off_bar = off_out + (kEpi == 2 ? kGemmAuxOutBytes : kGemmOutBytes);
  static_assert(off_bar + kGemmBarBytes + 1024 <= kGemmSmemBytes, "shared memory budget");
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem + off_bar);
  const uint32_t smem_base = smem_u32(smem);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t num_tiles = a.tiles_m * a.tiles_n;
  const uint32_t num_kb = (a.K + kGemmBK - 1) / kGemmBK;
  // work item w = tile * splits + split; a split owns k blocks [split * kb_per, ... + kb_per)
  const uint32_t num_work = num_tiles * a.splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->tmem_full[b]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[b]), kWide ? 16 : 8);
    }
    for (int b = 0; b < 4; ++b) mbar_init(smem_u32(&bars->res_full[b]), 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_d);
    if (a.flags & (kGemmResidual | kGemmGeluBwd)) prefetch_tmap(&tmap_r);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bars->tmem_base), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if constexpr (kWide) {
    // TMA / MMA / allocator / spare warps hand registers to the epilogue warps: the pool is 640 x 96 registers,
    // 128 x 56 + 512 x 104 fits
    if (warp < 4) reg_dealloc<56>();
  }
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        const uint32_t t = w / a.splits, sp = w % a.splits;
        const int32_t m0 = static_cast<int32_t>((t / a.tiles_n) * kGemmBM);
        const int32_t n0 = static_cast<int32_t>((t % a.tiles_n) * kGemmBN);
        const uint32_t kb0 = sp * a.kb_per_split;
        const uint32_t kb1 = min(num_kb, kb0 + a.kb_per_split);
        for (uint32_t kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t fbar = smem_u32(&bars->full[stage]);
          mbar_arrive_expect_tx(fbar, kGemmStageBytes);
          const uint32_t sa = smem_base + stage * kGemmStageBytes;
          const int32_t k0 = static_cast<int32_t>(kb * kGemmBK);
          if constexpr (kAMN) {
#pragma unroll
            for (int at = 0; at < kGemmBM / 64; ++at)
              tma_load_2d(sa + at * 8192, &tmap_a, fbar, m0 + at * 64, k0, kEvictNormal);
          } else {
            tma_load_2d(sa, &tmap_a, fbar, k0, m0, kEvictNormal);
          }
          if constexpr (kBMN) {
#pragma unroll
            for (int at = 0; at < kGemmBN / 64; ++at)
              tma_load_2d(sa + kGemmStageA + at * 8192, &tmap_b, fbar, n0 + at * 64, k0, kEvictLast);
          } else {
            tma_load_2d(sa + kGemmStageA, &tmap_b, fbar, k0, n0, kEvictLast);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (uint32_t w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
        const uint32_t sp = w % a.splits;
        const uint32_t kb0 = sp * a.kb_per_split;
        const uint32_t kb1 = min(num_kb, kb0 + a.kb_per_split);
        const uint32_t buf = it & 1;
        mbar_wait(smem_u32(&bars->tmem_empty[buf]), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kGemmBN;
        for (uint32_t kb = kb0; kb < kb1; ++kb) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kGemmStageBytes;
#pragma unroll
          for (int kk = 0; kk < kGemmBK / 16; ++kk) {
            const uint64_t adesc = kAMN ? smem_desc_sw128_mn(sa + kk * 2048, 8192, 1024)
                                        : smem_desc_sw128(sa) + static_cast<uint64_t>(kk * 2);
            const uint64_t bdesc = kBMN ? smem_desc_sw128_mn(sa + kGemmStageA + kk * 2048, 8192, 1024)
                                        : smem_desc_sw128(sa + kGemmStageA) + static_cast<uint64_t>(kk * 2);
            mma_f16_ss(d_tmem, adesc, bdesc, a.idesc, (kb != kb0 || kk != 0) ? 1u : 0u);
          }
          mma_commit(smem_u32(&bars->empty[stage]));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit(smem_u32(&bars->tmem_full[buf]));
      }
    }
  } else if (kEpi == 2 && warp >= 4) {
    // ===================================================== epilogue, sixteen warps, aux tile prefetched a tile ahead
    reg_alloc<104>();
    const uint32_t quad = warp & 3;          // TMEM lane quadrant this warp may read
    const uint32_t cq = (warp - 4) >> 2;     // column quarter of the tile: columns [cq * 64, cq * 64 + 64)
    const uint32_t row = quad * 32 + lane;   // tile row == TMEM lane
    const uint32_t stage_off = off_out + cq * (kGemmBM * 64 * 2);
    uint8_t* stage_ptr = smem + stage_off;
    const bool issuer = quad == 0 && lane == 0;
    const uint32_t bar_id = 1 + cq;
    const bool has_bias = (a.flags & kGemmBias) != 0;
    const bool has_gelu = (a.flags & kGemmGelu) != 0;
    const bool gelu_bwd = (a.flags & kGemmGeluBwd) != 0;
    const uint16_t* bias = static_cast<const uint16_t*>(a.bias);
    const uint32_t res_bar = smem_u32(&bars->res_full[cq]);
    const uint32_t col0 = cq * 64;
    uint32_t res_phase = 0;

    auto chunk_live = [&](uint32_t t) { return (t % a.tiles_n) * kGemmBN + col0 < a.N; };
    auto load_aux = [&](uint32_t t) {
      mbar_arrive_expect_tx(res_bar, kGemmBM * 64 * 2);
      tma_load_2d(smem_base + stage_off, &tmap_r, res_bar, static_cast<int32_t>((t % a.tiles_n) * kGemmBN + col0),
                  static_cast<int32_t>((t / a.tiles_n) * kGemmBM), kEvictNormal);
    };
    // first tile at or after t whose chunk of this column quarter exists (ragged N); >= num_work if none
    auto next_live = [&](uint32_t t) {
      while (t < num_work && !chunk_live(t)) t += gridDim.x;
      return t;
    };
    if (issuer) {
      const uint32_t ft = next_live(blockIdx.x);
      if (ft < num_work) load_aux(ft);
    }

    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < num_work; t += gridDim.x, ++it) {   // splits == 1: work item == tile
      const uint32_t buf = it & 1;
      const uint32_t m0 = (t / a.tiles_n) * kGemmBM;
      const uint32_t n0 = (t % a.tiles_n) * kGemmBN;
      if (has_bias) {
        const uint32_t e = threadIdx.x - 128;
        if (e < kGemmBN) bars->bias_stage[buf][e] = n0 + e < a.N ? bias[n0 + e] : static_cast<uint16_t>(0);
      }
      mbar_wait(smem_u32(&bars->tmem_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      if (has_bias) named_bar_sync(5, kGemmWideThreads - 128);
      const uint32_t gcol = n0 + col0;
      const bool live = gcol < a.N;
      const uint32_t t_addr = tmem_base + ((quad * 32) << 16) + buf * kGemmBN + col0;
      if (live) {
        mbar_wait(res_bar, res_phase);   // this chunk's aux box is in the staging box
        res_phase ^= 1;
      }
      uint32_t packed[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {   // 32 accumulators in registers at a time
        uint32_t v[32];
        if (live) {
          tmem_ld_32x32b_x32(t_addr + hh * 32, v);
          tmem_ld_wait();
        }
        if (hh == 1) {   // this warp has read everything it needs from the accumulator buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
        }
        if (!live) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {   // 8 columns per group == one 16-byte slot
          const int gg = hh * 4 + g;
          const uint32_t phys = (static_cast<uint32_t>(gg) ^ (row & 7u)) * 16u;
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g * 8 + j]);
          if (has_bias) {   // columns past N read the zeros staged above
            const uint4 bv = *reinterpret_cast<const uint4*>(&bars->bias_stage[buf][col0 + gg * 8]);
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
          // out-of-range rows / columns of the aux box were zero-filled by the TMA load
          const uint4 rv = *reinterpret_cast<const uint4*>(stage_ptr + row * 128u + phys);
          const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack2<kBf16>(rw[j]);
            if (gelu_bwd) {   // aux = pre-activation u: dU = dA * GeLU'(u)
              gelu_erf_grad_pair(f.x, f.y, x[2 * j], x[2 * j + 1]);
            } else {
              x[2 * j] += f.x;
              x[2 * j + 1] += f.y;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) packed[gg * 4 + j] = pack2<kBf16>(x[2 * j], x[2 * j + 1]);
        }
      }
      if (!live) continue;
      // the box is this chunk's: every thread overwrites exactly the 16-byte slots it has read
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
        *reinterpret_cast<uint4*>(stage_ptr + row * 128u + phys) =
            make_uint4(packed[g * 4], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
      }
      fence_proxy_async_smem();
      named_bar_sync(bar_id, 128);
      if (issuer) {
        tma_store_2d(&tmap_d, smem_base + stage_off, static_cast<int32_t>(gcol), static_cast<int32_t>(m0));
        tma_store_commit();
        const uint32_t nt = next_live(t + gridDim.x);   // aux of this quarter's next live chunk
        if (nt < num_work) {
          tma_store_wait_read<0>();
          load_aux(nt);
        }
      }
    }
    if (issuer) tma_store_wait<0>();
  } else if (kEpi == 1 && warp >= 4) {
    // ===================================================== epilogue, sixteen warps (write-only epilogues)
    reg_alloc<104>();
    const uint32_t quad = warp & 3;          // TMEM lane quadrant this warp may read
    const uint32_t cq = (warp - 4) >> 2;     // column quarter of the tile: columns [cq * 64, cq * 64 + 64)
    const uint32_t row = quad * 32 + lane;   // tile row == TMEM lane
    const uint32_t box_off = off_out + cq * (kGemmBM * 32 * 2);
    uint8_t* box_row = smem + box_off + row * 64u;
    const uint32_t swz = (row >> 1) & 3u;    // 64-byte swizzle: 16-byte slot index ^= address bits [7, 9)
    const bool issuer = quad == 0 && lane == 0;
    const uint32_t bar_id = 1 + cq;
    const bool has_bias = (a.flags & kGemmBias) != 0;
    const bool has_gelu = (a.flags & kGemmGelu) != 0;
    const bool store_pre = (a.flags & kGemmPreact) != 0;
    const uint16_t* bias = static_cast<const uint16_t*>(a.bias);
    const uint32_t col0 = cq * 64;

    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < num_work; t += gridDim.x, ++it) {   // splits == 1: work item == tile
      const uint32_t buf = it & 1;
      const uint32_t m0 = (t / a.tiles_n) * kGemmBM;
      const uint32_t n0 = (t % a.tiles_n) * kGemmBN;
      if (has_bias) {   // the first eight epilogue warps stage the tile's 256 bias values
        const uint32_t e = threadIdx.x - 128;
        if (e < kGemmBN) bars->bias_stage[buf][e] = n0 + e < a.N ? bias[n0 + e] : static_cast<uint16_t>(0);
      }
      mbar_wait(smem_u32(&bars->tmem_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      if (has_bias) named_bar_sync(5, kGemmWideThreads - 128);
      const uint32_t t_addr = tmem_base + ((quad * 32) << 16) + buf * kGemmBN + col0;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {       // the quarter's two 32-column halves
        const uint32_t gcol = n0 + col0 + hh * 32;
        const bool live = gcol < a.N;
        uint32_t v[32];
        if (live) {
          tmem_ld_32x32b_x32(t_addr + hh * 32, v);
          tmem_ld_wait();
        }
        if (hh == 1) {   // this warp has read everything it needs from the accumulator buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
        }
        if (!live) continue;
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        if (has_bias) {   // columns past N read the zeros staged above
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 bv = *reinterpret_cast<const uint4*>(&bars->bias_stage[buf][col0 + hh * 32 + g * 8]);
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack2<kBf16>(bw[j]);
              x[g * 8 + 2 * j] += f.x;
              x[g * 8 + 2 * j + 1] += f.y;
            }
          }
        }
        if (store_pre) {
          // training forward of the h -> 4h projection: the pre-activation (bias added) goes out first
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);       // the box's previous store has drained it
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(box_row + ((static_cast<uint32_t>(g) ^ swz) * 16u)) =
                make_uint4(pack2<kBf16>(x[g * 8], x[g * 8 + 1]), pack2<kBf16>(x[g * 8 + 2], x[g * 8 + 3]),
                           pack2<kBf16>(x[g * 8 + 4], x[g * 8 + 5]), pack2<kBf16>(x[g * 8 + 6], x[g * 8 + 7]));
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (issuer) {
            tma_store_2d(&tmap_p, smem_base + box_off, static_cast<int32_t>(gcol), static_cast<int32_t>(m0));
            tma_store_commit();
          }
        }
        if (has_gelu) {   // arithmetic first: the previous store drains the box underneath it
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_erf_pair(x[j], x[j + 1]);
        }
        if (issuer) tma_store_wait_read<0>();
        named_bar_sync(bar_id, 128);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(box_row + ((static_cast<uint32_t>(g) ^ swz) * 16u)) =
              make_uint4(pack2<kBf16>(x[g * 8], x[g * 8 + 1]), pack2<kBf16>(x[g * 8 + 2], x[g * 8 + 3]),
                         pack2<kBf16>(x[g * 8 + 4], x[g * 8 + 5]), pack2<kBf16>(x[g * 8 + 6], x[g * 8 + 7]));
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer) {
          tma_store_2d(&tmap_d, smem_base + box_off, static_cast<int32_t>(gcol), static_cast<int32_t>(m0));
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait<0>();
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const uint32_t quad = warp & 3;          // TMEM lane quadrant this warp may read
    const uint32_t hh = (warp - 4) >> 2;     // column half of the tile
    const uint32_t row = quad * 32 + lane;   // tile row == TMEM lane
    const uint32_t stage_off = off_out + hh * (kGemmBM * 64 * 2);
    uint8_t* stage_ptr = smem + stage_off;
    const bool issuer = (warp - 4) % 4 == 0 && lane == 0;
    const uint32_t bar_id = 1 + hh;
    const bool has_bias = (a.flags & kGemmBias) != 0;
    const bool has_gelu = (a.flags & kGemmGelu) != 0;
    const bool has_res = (a.flags & (kGemmResidual | kGemmGeluBwd)) != 0;   // an aux tile comes in by TMA
    const bool gelu_bwd = (a.flags & kGemmGeluBwd) != 0;
    const bool store_pre = (a.flags & kGemmPreact) != 0;
    const bool out_f32 = (a.flags & kGemmAccumF32) != 0;
    const uint16_t* bias = static_cast<const uint16_t*>(a.bias);
    const uint32_t res_bar = smem_u32(&bars->res_full[hh]);
    uint32_t res_phase = 0;

    // The residual tile of a chunk is fetched by TMA into the chunk's own staging buffer (coalesced,
    // asynchronous) and the result overwrites it in place.  `load_residual(t, ch)` is called by the
    // issuer thread as soon as the buffer is free, i.e. one chunk ahead of its use.
    auto chunk_live = [&](uint32_t w, uint32_t ch) {
      const uint32_t t = w / a.splits;
      return w < num_work && (t % a.tiles_n) * kGemmBN + hh * 128 + ch * 64 < a.N;
    };
    auto load_residual = [&](uint32_t w, uint32_t ch) {
      const uint32_t t = w / a.splits;
      mbar_arrive_expect_tx(res_bar, kGemmBM * 64 * 2);
      tma_load_2d(smem_base + stage_off, &tmap_r, res_bar,
                  static_cast<int32_t>((t % a.tiles_n) * kGemmBN + hh * 128 + ch * 64),
                  static_cast<int32_t>((t / a.tiles_n) * kGemmBM), kEvictNormal);
    };
    // First live chunk at or after (w, ch) in this column half's processing order (ragged N: the
    // chunks of the last column tile may be past the end for one half); w >= num_work if none.
    auto next_live_chunk = [&](uint32_t& w, uint32_t& ch) {
      while (w < num_work) {
        if (ch == 2) {
          w += gridDim.x;
          ch = 0;
          continue;
        }
        if (chunk_live(w, ch)) return;
        ++ch;
      }
    };
    if (has_res && !out_f32 && issuer) {
      uint32_t fw = blockIdx.x, fch = 0;
      next_live_chunk(fw, fch);
      if (fw < num_work) load_residual(fw, fch);
    }

    uint32_t it = 0;
    for (uint32_t w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const uint32_t t = w / a.splits;
      const uint32_t buf = it & 1;
      const uint32_t m0 = (t / a.tiles_n) * kGemmBM;
      const uint32_t n0 = (t % a.tiles_n) * kGemmBN;
      if (has_bias) {   // one 16-bit element per epilogue thread: columns n0 .. n0 + 255
        const uint32_t e = threadIdx.x - 128;
        bars->bias_stage[buf][e] = n0 + e < a.N ? bias[n0 + e] : static_cast<uint16_t>(0);
      }
      mbar_wait(smem_u32(&bars->tmem_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      if (has_bias) named_bar_sync(3, 256);   // slice visible; also keeps warps within one tile of each other
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
        if (ch == 1) {  // this warp has read everything it needs from the accumulator buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));
        }
        if (!live) continue;

        const uint32_t gcol = n0 + col0;
        if (out_f32) {
          // split-K / gradient accumulation: fp32 atomic adds straight into the output matrix
          if (m0 + row < a.M) {
            float* orow = a.out32 + static_cast<size_t>(m0 + row) * a.ldd32 + gcol;
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              if (gcol + g * 4 < a.N)
                atomicAdd(reinterpret_cast<float4*>(orow + g * 4),
                          make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]),
                                      __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3])));
            }
          }
          continue;
        }
        if (store_pre) {
          // training forward of the h -> 4h projection: also keep the pre-activation (bias added)
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
          mbar_wait(res_bar, res_phase);   // residual box is in the staging buffer
          res_phase ^= 1;
        }
        // The arithmetic runs BEFORE the wait for the staging buffer, so the previous chunk's TMA store
        // drains shared memory underneath it (it used to sit on the critical path of every chunk and
        // made the GeLU epilogue longer than the tile's MMAs).
        uint32_t packed[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {  // 8 columns per group == one 16-byte chunk
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
            const uint4 rv = *reinterpret_cast<const uint4*>(stage_ptr + row * 128u + phys);
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = unpack2<kBf16>(rw[j]);
              if (gelu_bwd) {   // aux = pre-activation u: dU = dA * GeLU'(u)
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
        if (!has_res) {
          // staging buffer must be free: the previous TMA store has finished reading it.  (With a
          // residual the buffer already is this chunk's: the residual box was loaded into it, and every
          // thread overwrites exactly the slots it has just read.)
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);
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
          if (has_res) {   // prefetch the residual of this group's next live chunk
            uint32_t nt = w, nch = ch + 1;
            next_live_chunk(nt, nch);
            if (nt < num_work) {
              tma_store_wait_read<0>();
              load_residual(nt, nch);
            }
          }
        }
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

template <bool kBf16, bool kAMN, bool kBMN, int kEpi>
cudaError_t prepare_one() {
  return cudaFuncSetAttribute(gemm_kernel<kBf16, kAMN, kBMN, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kGemmSmemBytes);
}

template <bool kBf16>
cudaError_t prepare_dtype() {
  cudaError_t e = prepare_one<kBf16, false, false, 0>();
  if (e == cudaSuccess) e = prepare_one<kBf16, false, true, 0>();
  if (e == cudaSuccess) e = prepare_one<kBf16, true, true, 0>();
  if (e == cudaSuccess) e = prepare_one<kBf16, false, false, 1>();
  if (e == cudaSuccess) e = prepare_one<kBf16, false, true, 1>();
  if (e == cudaSuccess) e = prepare_one<kBf16, false, false, 2>();
  if (e == cudaSuccess) e = prepare_one<kBf16, false, true, 2>();
  return e;
}

cudaError_t gemm_prepare() {
  cudaError_t e = prepare_dtype<true>();
  return e == cudaSuccess ? prepare_dtype<false>() : e;
}

namespace {

template <bool kBf16>
void launch_dtype(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const CUtensorMap& tmap_d, const CUtensorMap& tmap_r,
                  const CUtensorMap& tmap_p, const GemmArgs& args, bool a_mn, bool b_mn, int epi, int grid,
                  cudaStream_t stream) {
#define EMDR2_GEMM_LAUNCH(AMN, BMN, EPI)                                                                            \
  gemm_kernel<kBf16, AMN, BMN, EPI><<<grid, EPI != 0 ? kGemmWideThreads : kGemmThreads, kGemmSmemBytes, stream>>>(  \
      tmap_a, tmap_b, tmap_d, tmap_r, tmap_p, args)
  if (a_mn) EMDR2_GEMM_LAUNCH(true, true, 0);
  else if (b_mn && epi == 2) EMDR2_GEMM_LAUNCH(false, true, 2);
  else if (b_mn && epi == 1) EMDR2_GEMM_LAUNCH(false, true, 1);
  else if (b_mn) EMDR2_GEMM_LAUNCH(false, true, 0);
  else if (epi == 2) EMDR2_GEMM_LAUNCH(false, false, 2);
  else if (epi == 1) EMDR2_GEMM_LAUNCH(false, false, 1);
  else EMDR2_GEMM_LAUNCH(false, false, 0);
#undef EMDR2_GEMM_LAUNCH
}

}  // namespace

cudaError_t launch_gemm(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const CUtensorMap& tmap_d,
                        const CUtensorMap& tmap_r, const CUtensorMap& tmap_p, const GemmArgs& args,
                        bool bf16, bool a_mn, bool b_mn, int epi, int grid, cudaStream_t stream) {
  if (a_mn && !b_mn) return cudaErrorInvalidValue;   // not needed by any forward/backward product
  const bool has_aux = (args.flags & (kGemmResidual | kGemmGeluBwd)) != 0;
  if (epi != 0 && (a_mn || args.splits != 1 || (args.flags & kGemmAccumF32))) return cudaErrorInvalidValue;
  if ((epi == 1 && has_aux) || (epi == 2 && (!has_aux || (args.flags & kGemmPreact)))) return cudaErrorInvalidValue;
  if (bf16) launch_dtype<true>(tmap_a, tmap_b, tmap_d, tmap_r, tmap_p, args, a_mn, b_mn, epi, grid, stream);
  else launch_dtype<false>(tmap_a, tmap_b, tmap_d, tmap_r, tmap_p, args, a_mn, b_mn, epi, grid, stream);
  return cudaGetLastError();
}

}  // namespace emdr2
