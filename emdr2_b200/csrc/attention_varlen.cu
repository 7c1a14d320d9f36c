// Variable-length fused attention forward over TOKEN-PACKED activations, sm_100a only, head dim 64.
//
// The towers of the read step see 400 sequences per question batch whose lengths spread over
// 100-250 tokens (NQ passages).  The reference pads them to rectangles of 256 / 512 and masks
// (megatron/model/emdr2_model.py:118-120,148-149); the rectangular kernel of attention_persist.cu
// skips all-padding 128-blocks and the host buckets sequences by length, one launch per bucket.
// Here the activations are packed back to back — [T, h] with T = sum of the real lengths, no
// padding anywhere — and ONE launch per layer walks an explicit work list:
//
//   item = (128-query tile of one sequence, head): q_row0 / q_valid rows of the packed Q matrix,
//          k_row0 / k_len rows of the packed K/V matrices, where the tile's output and lse go.
//
// Self-attention lists every (sequence, query tile, head); FiD cross-attention lists (question,
// head, key RANGE) so that the 10 000+ keys of a question are cut into ranges that run in parallel
// and are merged by their log-sum-exp weights afterwards.  There are no masks: a key either belongs
// to the item's range or does not exist (probability exactly 0, which is what the reference's
// masked_fill(-10000) + softmax gives a padding key in fp32: exp(-10000 - max) == 0).
//
// The pipeline per item is the persistent kernel's (attention_persist.cu): S = Q.K^T by tcgen05 into
// TMEM, one thread per query row reads its 128 scores, P goes back to TMEM as 16-bit over the scores,
// O += P.V with the A operand in TMEM, lazy running-max rescale; Q double-buffered, K/V in a TMA
// ring that runs ahead across items, two CTAs per SM.  Tiles are fetched with 2-D tensor maps over
// the whole packed matrix: a tile that starts near the end of its sequence also brings in rows of
// the NEXT sequence (or zeros past the end of the matrix) — those key columns get probability 0 and
// those query rows are never stored (partial tiles store row by row, full tiles by TMA).
// Items are sorted by decreasing cost on the host and dealt round-robin to the resident CTAs.
//
// Scheduling inside a CTA (what distinguishes this kernel from attention_persist.cu): P has its OWN tensor-memory
// columns — S [0,128) fp32, P [128,192) 16-bit pairs, O [192,256) — so the score tile is free again as soon as the
// softmax warps have pulled it into registers (`s_free`), long before P exists.  The MMA thread therefore issues
// S(j+1) — of the same item or of the next one — right after `s_free(j)` and only then waits for P(j): by the time the
// softmax warps have finished block j their next score tile has been sitting in TMEM for ~1000 cycles, and they run
// block after block without waiting on the tensor core (ncu of the previous schedule: MUFU pipe 46 % busy, warps
// mostly stalled on barriers; the exp2 stream is what bounds head-dim-64 attention).  `pv_done` orders the two
// things that must not overtake P.V(j-1): the O rescale and the overwrite of P by block j.
#include "attention.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegInf = -__builtin_huge_valf();
constexpr float kLazyRescale = 8.0f;    // log2 units the running max may lag behind

struct VBars {
  uint64_t q_full[2];
  uint64_t q_empty[2];
  uint64_t kv_full[kAttnStages];
  uint64_t kv_empty[kAttnStages];
  uint64_t s_full;     // S(g) has landed in TMEM                          (tcgen05.commit)
  uint64_t s_free;     // all four softmax warps hold S(g) in registers    (4 arrivals)
  uint64_t p_full;     // P(g) is in TMEM                                  (4 arrivals)
  uint64_t pv_done;    // P.V(g) has completed: P may be overwritten, O rescaled (tcgen05.commit)
  uint64_t o_full;     // last P.V of the item has completed
  uint32_t tmem_base;
};
static_assert(sizeof(VBars) <= kAttnBarBytes, "barrier block too large");

constexpr uint32_t kOffQ = 0;                                            // 2 x 16 KiB
constexpr uint32_t kOffKV = 2 * kAttnTileBytes;                          // ring of (K, V) pairs
constexpr uint32_t kOffO = kOffKV + kAttnStages * 2 * kAttnTileBytes;    // output staging
constexpr uint32_t kOffBar = kOffO + kAttnTileBytes;
constexpr int kVarlenSmemBytes = kOffBar + kAttnBarBytes;
static_assert(kVarlenSmemBytes <= 113 * 1024, "two CTAs per SM must fit");

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Debug build only (-DEMDR2_VARLEN_TRACE, tools/gpu_trace_varlen.py): clock stamps of CTA 0's first softmax warp and
// of its MMA thread at the hand-over points of every key block, read back with emdr2_varlen_trace_read.
#ifdef EMDR2_VARLEN_TRACE
constexpr int kTraceBlocks = 64, kTraceSlots = 8;
__device__ long long g_varlen_trace[2][kTraceBlocks][kTraceSlots];
#define VTRACE(role, blk, slot)                                                                   \
  do {                                                                                            \
    if (blockIdx.x == 0 && (blk) < kTraceBlocks) g_varlen_trace[role][blk][slot] = clock64();     \
  } while (0)
#else
#define VTRACE(role, blk, slot) \
  do {                          \
  } while (0)
#endif

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

__device__ __forceinline__ AttnVarlenItem load_item(const AttnVarlenItem* items, uint32_t idx) {
  const int4* p = reinterpret_cast<const int4*>(items + idx);
  const int4 a = __ldg(p), b = __ldg(p + 1);
  AttnVarlenItem it;
  it.q_row0 = a.x;
  it.q_valid = a.y;
  it.k_row0 = a.z;
  it.k_len = a.w;
  it.head = b.x;
  it.o_row0 = b.y;
  it.lse_idx0 = b.z;
  it.reserved = b.w;
  return it;
}

template <bool kBf16>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_varlen_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                        const AttnVarlenArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SW128 tiles need 1024-B alignment
  VBars* bars = reinterpret_cast<VBars*>(smem + kOffBar);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;

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
    mbar_init(smem_u32(&bars->s_free), 4);
    mbar_init(smem_u32(&bars->p_full), 4);
    mbar_init(smem_u32(&bars->pv_done), 1);
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
  const uint32_t tmem_p = tmem_base + kAttnBK;             // 64 columns of 16-bit pairs
  const uint32_t tmem_o = tmem_base + kAttnBK + 64;        // 64 fp32 columns

  if (warp < 4) {
    reg_dealloc<40>();   // warpgroup 0 (TMA, MMA, allocator, spare) hands its registers to the softmax warps
    if (warp == 0 && lane == 0) {
      // ===================================================== TMA producer
      uint32_t stage = 0, phase = 0, n = 0;
      for (uint32_t idx = blockIdx.x; idx < a.n_items; idx += gridDim.x, ++n) {
        const AttnVarlenItem it = load_item(a.items, idx);
        const int32_t col_h = it.head * kAttnHeadDim;
        const uint32_t buf = n & 1;
        mbar_wait_backoff<200>(smem_u32(&bars->q_empty[buf]), ((n >> 1) & 1) ^ 1);
        const uint32_t qbar = smem_u32(&bars->q_full[buf]);
        mbar_arrive_expect_tx(qbar, kAttnTileBytes);
        tma_load_2d(smem_base + kOffQ + buf * kAttnTileBytes, &tmap_q, qbar, col_h, it.q_row0, kEvictNormal);
        const uint32_t nblk = (static_cast<uint32_t>(it.k_len) + kAttnBK - 1) / kAttnBK;
        for (uint32_t j = 0; j < nblk; ++j) {
          mbar_wait_backoff<100>(smem_u32(&bars->kv_empty[stage]), phase ^ 1);
          const uint32_t fbar = smem_u32(&bars->kv_full[stage]);
          mbar_arrive_expect_tx(fbar, 2 * kAttnTileBytes);
          const uint32_t dst = smem_base + kOffKV + stage * 2 * kAttnTileBytes;
          const int32_t krow = it.k_row0 + static_cast<int32_t>(j * kAttnBK);
          tma_load_2d(dst, &tmap_k, fbar, col_h, krow, kEvictLast);
          tma_load_2d(dst + kAttnTileBytes, &tmap_v, fbar, col_h, krow, kEvictLast);
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================================================== MMA issuer
      // Two cursors over the same (item, key block) sequence: `sc` issues score products one block AHEAD of `pc`,
      // which issues the P.V products.  Order of issue: S(0); then per block g: S(g+1) once S(g) is in registers,
      // P.V(g) once P(g) is in TMEM.
      struct Cursor {
        uint32_t idx, n, j, nblk;
        bool ok;
      };
      auto load = [&](Cursor& c) {
        c.ok = c.idx < a.n_items;
        if (c.ok) c.nblk = (static_cast<uint32_t>(load_item(a.items, c.idx).k_len) + kAttnBK - 1) / kAttnBK;
      };
      auto advance = [&](Cursor& c) {
        if (++c.j == c.nblk) {
          c.idx += gridDim.x;
          ++c.n;
          c.j = 0;
          load(c);
        }
      };
      Cursor sc{blockIdx.x, 0, 0, 0, false}, pc{blockIdx.x, 0, 0, 0, false};
      load(sc);
      load(pc);
      uint32_t s_stage = 0, s_phase = 0;   // K/V ring position of the next S product
      uint32_t pv_stage = 0;               // ... and of the next P.V product
      auto issue_s = [&]() {
        const uint32_t buf = sc.n & 1;
        if (sc.j == 0) {
          mbar_wait(smem_u32(&bars->q_full[buf]), (sc.n >> 1) & 1);
          tc_fence_after();
        }
        mbar_wait(smem_u32(&bars->kv_full[s_stage]), s_phase);
        tc_fence_after();
        const uint64_t qdesc = smem_desc_sw128(smem_base + kOffQ + buf * kAttnTileBytes);
        const uint64_t kdesc = smem_desc_sw128(smem_base + kOffKV + s_stage * 2 * kAttnTileBytes);
#pragma unroll
        for (int kk = 0; kk < kAttnHeadDim / 16; ++kk)
          mma_f16_ss(tmem_base, qdesc + static_cast<uint64_t>(kk * 2), kdesc + static_cast<uint64_t>(kk * 2),
                     a.idesc_s, kk != 0 ? 1u : 0u);
        mma_commit(smem_u32(&bars->s_full));
        if (sc.j + 1 == sc.nblk) mma_commit(smem_u32(&bars->q_empty[buf]));   // Q is only read by the S products
        if (++s_stage == kAttnStages) {
          s_stage = 0;
          s_phase ^= 1;
        }
        advance(sc);
      };
      if (sc.ok) issue_s();
      for (uint32_t g = 0; pc.ok; ++g) {
        VTRACE(1, g, 0);
        if (sc.ok && a.sched == 0) {
          mbar_wait(smem_u32(&bars->s_free), g & 1);     // S(g) is in the softmax warps' registers
          tc_fence_after();
          VTRACE(1, g, 1);
          issue_s();                                     // S(g+1)
        }
        VTRACE(1, g, 2);
        mbar_wait(smem_u32(&bars->p_full), g & 1);
        tc_fence_after();
        VTRACE(1, g, 3);
        const uint32_t vbase = smem_base + kOffKV + pv_stage * 2 * kAttnTileBytes + kAttnTileBytes;
#pragma unroll
        for (int ks = 0; ks < kAttnBK / 16; ++ks) {
          // A = P in TMEM: row = lane, 16 keys = 8 packed 32-bit columns; B = V MN-major, 16 keys =
          // two 8-row groups of 1024 B
          const uint64_t vdesc = smem_desc_sw128_mn(vbase + ks * 2048, 1024, 1024);
          mma_f16_ts(tmem_o, tmem_p + ks * 8, vdesc, a.idesc_o, (pc.j != 0 || ks != 0) ? 1u : 0u);
        }
        mma_commit(smem_u32(&bars->kv_empty[pv_stage]));
        mma_commit(smem_u32(&bars->pv_done));
        if (pc.j + 1 == pc.nblk) mma_commit(smem_u32(&bars->o_full));
        if (++pv_stage == kAttnStages) pv_stage = 0;
        VTRACE(1, g, 4);
        advance(pc);
        if (sc.ok && a.sched != 0) issue_s();            // conservative order (debug): S(g+1) behind P.V(g)
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
    uint32_t blk = 0, n = 0;
    bool store_pending = false;
    uint16_t* out_base = static_cast<uint16_t*>(a.out);

    for (uint32_t idx = blockIdx.x; idx < a.n_items; idx += gridDim.x, ++n) {
      const AttnVarlenItem it = load_item(a.items, idx);
      const uint32_t q_valid = static_cast<uint32_t>(it.q_valid);
      const uint32_t k_len = static_cast<uint32_t>(it.k_len);
      const uint32_t nblk = (k_len + kAttnBK - 1) / kAttnBK;
      const bool warp_active = quad * 32 < q_valid;      // warps whose 32 rows are all past the tile's end idle
      const bool row_active = row < q_valid;
      uint4 out[8];
      float m_run = kNegInf;
      float l_run = 0.f;

      for (uint32_t j = 0; j < nblk; ++j, ++blk) {
        const uint32_t valid = min(static_cast<uint32_t>(kAttnBK), k_len - j * kAttnBK);
        if (threadIdx.x == 128) VTRACE(0, blk, 0);
        mbar_wait(smem_u32(&bars->s_full), blk & 1);
        tc_fence_after();
        if (threadIdx.x == 128) VTRACE(0, blk, 1);
        if (warp_active) {
          uint32_t t[kAttnBK];   // scores as raw fp32 bits (tcgen05.ld output registers)
          const uint32_t s_addr = tmem_base + lane_tmem;
          tmem_ld_32x32b_x32(s_addr, *reinterpret_cast<uint32_t(*)[32]>(&t[0]));
          tmem_ld_32x32b_x32(s_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&t[32]));
          tmem_ld_32x32b_x32(s_addr + 64, *reinterpret_cast<uint32_t(*)[32]>(&t[64]));
          tmem_ld_32x32b_x32(s_addr + 96, *reinterpret_cast<uint32_t(*)[32]>(&t[96]));
          tmem_ld_wait();
          if (threadIdx.x == 128) VTRACE(0, blk, 2);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->s_free));   // the tensor core may overwrite S now
          // ---- block maximum.  A full block (every block but a range's last): raw maximum, one scale.
          // The last block: 32-key chunks inside the range keep their raw accumulators, chunks past its
          // end contribute nothing (multiplier 0, bias -inf), and the one chunk the end falls into is
          // rewritten element by element.
          float mx;
          float mul_c[4], bias_c[4];
          if (valid == kAttnBK) {
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
              const uint32_t c0 = c * 32;
              float m;
              if (c0 >= valid) {                       // wholly past the end of the key range
                m = kNegInf;
                mul_c[c] = 0.f;
                bias_c[c] = kNegInf;
              } else if (c0 + 32 <= valid) {           // wholly inside
                float r0 = __uint_as_float(t[c0]), r1 = __uint_as_float(t[c0 + 1]);
#pragma unroll
                for (int i = 2; i < 32; i += 2) {
                  r0 = fmaxf(r0, __uint_as_float(t[c0 + i]));
                  r1 = fmaxf(r1, __uint_as_float(t[c0 + i + 1]));
                }
                m = fmaxf(r0, r1) * a.scale_log2;
                mul_c[c] = a.scale_log2;
                bias_c[c] = 0.f;
              } else {                                 // the end falls into this chunk
                m = kNegInf;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  float x = __uint_as_float(t[c0 + i]) * a.scale_log2;
                  x = (c0 + i < valid) ? x : kNegInf;
                  t[c0 + i] = __float_as_uint(x);
                  m = fmaxf(m, x);
                }
                mul_c[c] = 1.0f;
                bias_c[c] = 0.f;
              }
              mx = fmaxf(mx, m);
            }
          }
          // P.V of the previous block (of this item or the one before) must have completed before O is rescaled
          // or P overwritten; it was issued ~a block ago, so this wait is normally free
          if (threadIdx.x == 128) VTRACE(0, blk, 3);
          if (blk > 0) {
            mbar_wait(smem_u32(&bars->pv_done), (blk - 1) & 1);
            tc_fence_after();
          }
          if (threadIdx.x == 128) VTRACE(0, blk, 4);
          // ---- running maximum: raised only when the block exceeds it by more than 2^8
          const bool first = j == 0;
          const bool raise = first || mx > m_run + kLazyRescale;
          float alpha = 1.0f;
          if (raise) {
            alpha = ex2(m_run - mx);   // 0 for the first block (m_run = -inf)
            m_run = mx;
          }
          l_run *= alpha;
          if (!first && __any_sync(kFull, raise)) {
            // O *= alpha in TMEM.  Warp-uniform branch: tcgen05.ld/st are .sync.aligned.
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
          // ---- P = exp2(t * mul + bias - m_run) -> 16-bit pairs -> TMEM columns [0, 64)
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int c = 0; c < kAttnBK / 32; ++c) {
            const float mul = mul_c[c];
            const float negm = bias_c[c] - m_run;
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2(fmaf(__uint_as_float(t[c * 32 + 2 * i]), mul, negm));
              const float p1 = ex2(fmaf(__uint_as_float(t[c * 32 + 2 * i + 1]), mul, negm));
              sum0 += p0;
              sum1 += p1;
              w[i] = pack2<kBf16>(p0, p1);
            }
            tmem_st_32x32b_x16(tmem_p + lane_tmem + c * 16, w);
          }
          l_run += sum0 + sum1;
          tmem_st_wait();
          if (threadIdx.x == 128) VTRACE(0, blk, 5);
        } else {
          // a warp whose 32 rows are all past the end of the tile: its rows of P / O are never stored (the tensor
          // core computes them from whatever the lanes hold; rows are independent) — it only keeps the barriers going
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->s_free));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->p_full));
        // An idle warp has nothing to do until the next score tile, which now lands BEFORE the working warps have
        // finished this block: without this wait it would run a block ahead and its next arrival would be counted
        // into THIS block's p_full phase, releasing P.V before P is written.
        if (!warp_active) mbar_wait(smem_u32(&bars->p_full), blk & 1);
      }

      // ---- item epilogue: O / l -> 16-bit row
      if (threadIdx.x == 128) VTRACE(0, blk - 1, 6);
      mbar_wait(smem_u32(&bars->o_full), n & 1);
      tc_fence_after();
      if (threadIdx.x == 128) VTRACE(0, blk - 1, 7);
      if (warp_active) {
        const float inv_l = 1.0f / l_run;
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
        if (a.lse && row_active) a.lse[static_cast<size_t>(it.lse_idx0) + row] = (m_run + log2f(l_run)) * kLn2;
      }

      if (q_valid == kAttnBQ) {
        // ---- full tile: stage the 128 x 64 block and hand it to TMA
        if (store_thread && store_pending) tma_store_wait_read<0>();   // previous tile has left smem
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t phys = (static_cast<uint32_t>(g) ^ (row & 7u)) * 16u;
          *reinterpret_cast<uint4*>(o_row + phys) = out[g];
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (store_thread) {
          tma_store_2d(&tmap_o, smem_base + kOffO, it.head * kAttnHeadDim, it.o_row0);
          tma_store_commit();
          store_pending = true;
        }
      } else if (row_active) {
        // ---- partial tile (the end of a sequence): every thread writes its own 128-byte row; rows past
        // q_valid belong to the next sequence and must not be touched
        uint4* dst = reinterpret_cast<uint4*>(out_base + (static_cast<size_t>(it.o_row0) + row) * a.ldo +
                                              it.head * kAttnHeadDim);
#pragma unroll
        for (int g = 0; g < 8; ++g) dst[g] = out[g];
      }
    }
    if (store_thread && store_pending) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kAttnTmemCols);
}

}  // namespace

#ifdef EMDR2_VARLEN_TRACE
extern "C" __attribute__((visibility("default"))) int emdr2_varlen_trace_read(long long* out, int n) {
  const size_t bytes = sizeof(long long) * static_cast<size_t>(n < 2 * kTraceBlocks * kTraceSlots ? n : 2 * kTraceBlocks * kTraceSlots);
  return static_cast<int>(cudaMemcpyFromSymbol(out, g_varlen_trace, bytes));
}
#endif

cudaError_t attention_varlen_prepare() {
  cudaError_t e = cudaFuncSetAttribute(attention_varlen_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kVarlenSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention_varlen_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              kVarlenSmemBytes);
}

void launch_attention_varlen(const CUtensorMap& tmap_q, const CUtensorMap& tmap_k, const CUtensorMap& tmap_v,
                             const CUtensorMap& tmap_o, const AttnVarlenArgs& args, bool bf16, int sm_count,
                             cudaStream_t stream) {
  if (args.n_items == 0) return;
  const uint32_t slots = 2u * static_cast<uint32_t>(sm_count);
  const dim3 grid(args.n_items < slots ? args.n_items : slots);
  if (bf16)
    attention_varlen_kernel<true><<<grid, kAttnThreads, kVarlenSmemBytes, stream>>>(tmap_q, tmap_k, tmap_v, tmap_o, args);
  else
    attention_varlen_kernel<false><<<grid, kAttnThreads, kVarlenSmemBytes, stream>>>(tmap_q, tmap_k, tmap_v, tmap_o, args);
}

}  // namespace emdr2
