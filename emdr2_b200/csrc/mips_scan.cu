// See mips_scan.cuh for the design. sm_100a only.
#include "mips_scan.cuh"
#include "ptx.cuh"

namespace emdr2 {
using namespace ptx;

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kNegInfBits = 0xff800000u;

struct Bars {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t q_full;
  uint32_t tmem_base;
  volatile uint32_t done;
};
static_assert(sizeof(Bars) <= kBarBytes, "barrier block too large");

__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// Sort 128 keys held 4 per lane (element index = lane*4 + r) into descending order.
__device__ __forceinline__ void warp_bitonic_desc128(uint64_t (&key)[4], uint32_t lane) {
#pragma unroll
  for (int size = 2; size <= 128; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      if (stride >= 4) {
        const int ls = stride >> 2;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint64_t other = __shfl_xor_sync(kFull, key[r], ls);
          const int i = static_cast<int>(lane) * 4 + r;
          const bool desc_block = (i & size) == 0;
          const bool lower = (i & stride) == 0;
          const uint64_t mx = umax64(key[r], other), mn = umin64(key[r], other);
          key[r] = (lower == desc_block) ? mx : mn;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int p = r ^ stride;
          if (p > r) {
            const int i = static_cast<int>(lane) * 4 + r;
            const bool desc_block = (i & size) == 0;
            const uint64_t a = key[r], b = key[p];
            const uint64_t mx = umax64(a, b), mn = umin64(a, b);
            key[r] = desc_block ? mx : mn;
            key[p] = desc_block ? mn : mx;
          }
        }
      }
    }
  }
}

// Candidate entry in shared memory: .x = fp32 score bits, .y = shard-local row.
// Sort key: larger is better under (score desc, row asc).
__device__ __forceinline__ uint64_t cand_key(uint2 e) {
  return (static_cast<uint64_t>(f32_to_ordered(e.x)) << 32) | static_cast<uint64_t>(~e.y);
}
__device__ __forceinline__ uint2 key_cand(uint64_t key) {
  uint2 e;
  e.x = ordered_to_f32(static_cast<uint32_t>(key >> 32));
  e.y = ~static_cast<uint32_t>(key);
  return e;
}

// Warp-cooperative compaction of one query's candidate list: sort the n (<= kCap) entries by
// (score desc, row asc), keep the best min(n, k) in place (sorted).  Every lane returns the same
// values: new count, and (if n >= k) the k-th best entry.
__device__ __forceinline__ void compact_list(uint2* buf, int n, int k, uint32_t lane, int& new_n,
                                             uint2& kth, bool& has_kth) {
  uint64_t key[4];
  uint2 e[4];
  if (lane * 4 < static_cast<uint32_t>(kCap)) {
    const uint4* b4 = reinterpret_cast<const uint4*>(buf) + lane * 2;
    const uint4 lo = b4[0], hi = b4[1];
    e[0] = make_uint2(lo.x, lo.y);
    e[1] = make_uint2(lo.z, lo.w);
    e[2] = make_uint2(hi.x, hi.y);
    e[3] = make_uint2(hi.z, hi.w);
  } else {
    e[0] = e[1] = e[2] = e[3] = make_uint2(0u, 0u);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = static_cast<int>(lane) * 4 + r;
    key[r] = (i < n) ? cand_key(e[r]) : 0ull;
  }
  __syncwarp();
  warp_bitonic_desc128(key, lane);
  new_n = n < k ? n : k;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = static_cast<int>(lane) * 4 + r;
    if (i < new_n) buf[i] = key_cand(key[r]);
  }
  has_kth = n >= k;
  const int kr = (k - 1) & 3;
  uint64_t sel = key[0];
  sel = kr == 1 ? key[1] : sel;
  sel = kr == 2 ? key[2] : sel;
  sel = kr == 3 ? key[3] : sel;
  sel = __shfl_sync(kFull, sel, (k - 1) >> 2);
  kth = key_cand(sel);
  __syncwarp();
}

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

}  // namespace

__global__ void __launch_bounds__(kScanThreads, 1)
mips_scan_kernel(const __grid_constant__ CUtensorMap tmap_q,
                 const __grid_constant__ CUtensorMap tmap_e, const ScanArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t off_stage = a.num_kb * kQBlockBytes;
  const uint32_t off_cand = off_stage + a.num_stages * kStageBytes;
  const uint32_t off_bar = off_cand + kCandBytes;
  Bars* bars = reinterpret_cast<Bars*>(smem + off_bar);
  uint2* cand = reinterpret_cast<uint2*>(smem + off_cand);
  const uint32_t smem_base = smem_u32(smem);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t cta = blockIdx.x;
  const uint32_t G = gridDim.x;
  const bool share = (a.flags & kFlagShare) != 0;
  const bool probe = (a.flags & kFlagProbe) != 0;

  // Tiles cta, cta+G, ... ; with probing the first tile is visited twice (probe pass, then for real).
  const uint32_t my_tiles = (a.num_tiles > cta) ? (a.num_tiles - cta + G - 1) / G : 0;
  const uint32_t num_iters = my_tiles + ((probe && my_tiles > 0) ? 1u : 0u);
  const uint32_t probe_off = probe ? 1u : 0u;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < a.num_stages; ++s) {
      mbar_init(smem_u32(&bars->full[s]), 1);
      mbar_init(smem_u32(&bars->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bars->tmem_full[b]), 1);
      mbar_init(smem_u32(&bars->tmem_empty[b]), 4);
    }
    mbar_init(smem_u32(&bars->q_full), 1);
    bars->done = 0;
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_e);
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer (one lane)
    if (lane == 0 && num_iters > 0) {
      const uint32_t qbar = smem_u32(&bars->q_full);
      mbar_arrive_expect_tx(qbar, a.num_kb * kQBlockBytes);
      for (uint32_t kb = 0; kb < a.num_kb; ++kb)
        tma_load_2d(smem_base + kb * kQBlockBytes, &tmap_q, qbar, static_cast<int32_t>(kb * kBlockK),
                    0, kEvictLast);
      uint32_t stage = 0, phase = 0;
      for (uint32_t it = 0; it < num_iters; ++it) {
        const uint32_t tile = cta + (it > probe_off ? it - probe_off : 0) * G;
        const int32_t row0 = static_cast<int32_t>(tile * kTileN);
        // The probed tile is read again right away: keep it in L2; everything else streams.
        const uint64_t policy = (probe && it == 0) ? kEvictLast : kEvictFirst;
        for (uint32_t kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(smem_u32(&bars->empty[stage]), phase ^ 1);
          const uint32_t fbar = smem_u32(&bars->full[stage]);
          mbar_arrive_expect_tx(fbar, kStageBytes);
          tma_load_2d(smem_base + off_stage + stage * kStageBytes, &tmap_e, fbar,
                      static_cast<int32_t>(kb * kBlockK), row0, policy);
          if (++stage == a.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one lane)
    if (lane == 0 && num_iters > 0) {
      mbar_wait(smem_u32(&bars->q_full), 0);
      tc_fence_after();
      uint32_t stage = 0, phase = 0;
      for (uint32_t it = 0; it < num_iters; ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(smem_u32(&bars->tmem_empty[buf]), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kTileN;
        for (uint32_t kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(smem_u32(&bars->full[stage]), phase);
          tc_fence_after();
          const uint64_t adesc = smem_desc_sw128(smem_base + kb * kQBlockBytes);
          const uint64_t bdesc = smem_desc_sw128(smem_base + off_stage + stage * kStageBytes);
#pragma unroll
          for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
            // advance both descriptors by 16 elements = 32 B inside the 128-B swizzle row
            mma_f16_ss(d_tmem, adesc + static_cast<uint64_t>(kk * 2),
                       bdesc + static_cast<uint64_t>(kk * 2), a.idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          mma_commit(smem_u32(&bars->empty[stage]));
          if (++stage == a.num_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit(smem_u32(&bars->tmem_full[buf]));
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================================================== shared-threshold service warps
    // Recomputes, for the queries this CTA serves, the k-th largest of the per-CTA running maxima
    // and publishes it (monotone atomicMax).  Needs at least k CTAs.
    if (share && G >= a.k && num_iters > 0) {
      constexpr int kPer = kMaxCtas / 32;
      const uint32_t per = (G + 31) >> 5;
      // warp 2 serves query (cta mod nq), warp 3 a query half the batch away: with 148 CTAs and
      // 64 queries every query has 4-5 independent servers polling out of phase.
      const uint32_t q_first = (cta + (warp == 3 ? (a.nq + 1) / 2 : 0u)) % a.nq;
      while (bars->done == 0) {
        for (uint32_t q = q_first; q < a.nq; q += G) {
          uint32_t v[kPer];
#pragma unroll
          for (int i = 0; i < kPer; ++i) {
            const uint32_t idx = lane + 32 * i;
            uint64_t x = 0;
            if (idx < G) x = ld_relaxed_u64(a.gmax + static_cast<size_t>(q) * kMaxCtas + idx);
            v[i] = (static_cast<uint32_t>(x >> 32) == a.epoch) ? static_cast<uint32_t>(x) : 0u;
          }
          // k-th largest by bisection on the ordered-uint key space: the largest t with
          // count(v >= t) >= k.  32 steps of kPer ballots each.
          uint32_t ans = 0;
#pragma unroll 1
          for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand_t = ans | (1u << bit);
            uint32_t c = 0;
#pragma unroll
            for (int i = 0; i < kPer; ++i)
              if (static_cast<uint32_t>(i) < per) c += __popc(__ballot_sync(kFull, v[i] >= cand_t));
            if (c >= a.k) ans = cand_t;
          }
          if (ans != 0u && lane == 0)
            atomicMax(reinterpret_cast<unsigned long long*>(a.gthr + q),
                      (static_cast<unsigned long long>(a.epoch) << 32) | ans);
        }
        __nanosleep(200);
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: TMEM -> threshold filter -> top-k lists
    const uint32_t ew = warp - 4;  // TMEM lane quadrant
    const uint32_t q = ew * 16 + lane;
    const bool active = lane < 16 && q < a.nq;
    const int k = static_cast<int>(a.k);
    uint2* my = cand + static_cast<size_t>(ew * 16 + (lane & 15)) * kCandStride;
    const bool use_share = share && G >= a.k;

    int cnt = 0;
    float loc_s = __uint_as_float(kNegInfBits);  // local k-th best score (valid once cnt reached k)
    uint32_t loc_row = 0xffffffffu;
    float g_s = __uint_as_float(kNegInfBits);    // shared lower bound
    float run_max = __uint_as_float(kNegInfBits);
    float pub_max = __uint_as_float(kNegInfBits);
    unsigned long long n_app = 0, n_cmp = 0, wait_ns = 0;

    for (uint32_t it = 0; it < num_iters; ++it) {
      const uint32_t buf = it & 1;
      const bool is_probe = probe && it == 0;
      const uint32_t tile = cta + (it > probe_off ? it - probe_off : 0) * G;
      const uint32_t row0 = tile * kTileN;
      const uint32_t valid = min(static_cast<uint32_t>(kTileN), a.n_rows - row0);

      if (use_share && active && !is_probe) {
        const uint64_t x = ld_relaxed_u64(a.gthr + q);
        if (static_cast<uint32_t>(x >> 32) == a.epoch)
          g_s = fmaxf(g_s, __uint_as_float(ordered_to_f32(static_cast<uint32_t>(x))));
      }
      float T = fmaxf(loc_s, g_s);
      uint32_t tie_row = (g_s > loc_s) ? 0xffffffffu : loc_row;

      mbar_wait(smem_u32(&bars->tmem_full[buf]), (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((ew * 32) << 16) + buf * kTileN;

#pragma unroll 1
      for (uint32_t c = 0; c < kTileN / 32; ++c) {
        if (c * 32 >= valid) break;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_addr + c * 32, v);
        tmem_ld_wait();
        const uint32_t col0 = c * 32;
        if (col0 + 32 > valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j >= valid) v[j] = kNegInfBits;
        }
        float mx = __uint_as_float(v[0]);
#pragma unroll
        for (int j = 1; j + 1 < 32; j += 2)
          mx = max3(mx, __uint_as_float(v[j]), __uint_as_float(v[j + 1]));
        mx = fmaxf(mx, __uint_as_float(v[31]));
        run_max = fmaxf(run_max, mx);
        if (is_probe) continue;
        const bool hit = active && (mx >= T);
        if (!__any_sync(kFull, hit)) continue;

        // ---- slow path: at least one query of this warp has a passing score in this chunk
        const uint32_t rbase = row0 + col0;
        uint32_t pm = 0;
        if (hit) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]);
            const bool p = (x > T) || (x == T && (rbase + j) < tie_row);
            pm |= (p ? 1u : 0u) << j;
          }
          if (col0 + 32 > valid) pm &= (valid - col0 >= 32) ? kFull : ((1u << (valid - col0)) - 1u);
        }
        uint32_t om = __ballot_sync(kFull, cnt + __popc(pm) > kCap);
        bool compacted = false;
        while (om) {
          const int L = __ffs(om) - 1;
          om &= om - 1;
          const int n = __shfl_sync(kFull, cnt, L);
          int new_n;
          uint2 kth;
          bool has_kth;
          compact_list(cand + static_cast<size_t>(ew * 16 + L) * kCandStride, n, k, lane, new_n,
                       kth, has_kth);
          if (static_cast<int>(lane) == L) {
            cnt = new_n;
            if (has_kth) {
              loc_s = __uint_as_float(kth.x);
              loc_row = kth.y;
            }
            compacted = true;
            ++n_cmp;
          }
        }
        if (compacted) {
          T = fmaxf(loc_s, g_s);
          tie_row = (g_s > loc_s) ? 0xffffffffu : loc_row;
          uint32_t pm2 = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]);
            const bool p = (x > T) || (x == T && (rbase + j) < tie_row);
            pm2 |= (p ? 1u : 0u) << j;
          }
          pm &= pm2;
        }
        if (pm) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if ((pm >> j) & 1u) {
              my[cnt] = make_uint2(v[j], rbase + j);
              ++cnt;
            }
          }
          n_app += __popc(pm);
        }
      }

      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->tmem_empty[buf]));

      // publish this CTA's running maximum for the query (valid: it is the score of a real row)
      if (use_share && active && run_max > pub_max) {
        pub_max = run_max;
        st_relaxed_u64(a.gmax + static_cast<size_t>(q) * kMaxCtas + cta,
                       (static_cast<uint64_t>(a.epoch) << 32) |
                           f32_to_ordered(__float_as_uint(run_max)));
      }
      if (is_probe && use_share) {
        // bounded wait for the first shared threshold so that the real pass starts filtered
        const uint64_t t0 = globaltimer_ns();
        bool ok = !active;
        while (true) {
          if (!ok) {
            const uint64_t x = ld_relaxed_u64(a.gthr + q);
            if (static_cast<uint32_t>(x >> 32) == a.epoch) {
              g_s = fmaxf(g_s, __uint_as_float(ordered_to_f32(static_cast<uint32_t>(x))));
              ok = true;
            }
          }
          const bool expired = globaltimer_ns() - t0 > a.probe_timeout_ns;
          if (__all_sync(kFull, ok) || __any_sync(kFull, expired)) break;
          __nanosleep(100);
        }
        wait_ns += globaltimer_ns() - t0;
      }
    }

    // ---- final: append this CTA's surviving candidates to the per-query pools in global memory.
    // Lists longer than k are cut to their best k first (rare: the shared bound keeps them short);
    // the merge kernel imposes the total order, so nothing needs sorting here.
    if (num_iters > 0) {
      uint32_t om = __ballot_sync(kFull, active && cnt > k);
      while (om) {
        const int L = __ffs(om) - 1;
        om &= om - 1;
        const int n = __shfl_sync(kFull, cnt, L);
        int new_n;
        uint2 kth;
        bool has_kth;
        compact_list(cand + static_cast<size_t>(ew * 16 + L) * kCandStride, n, k, lane, new_n, kth,
                     has_kth);
        if (static_cast<int>(lane) == L) cnt = new_n;
      }
      if (active && cnt > 0) {
        const uint32_t slot0 = atomicAdd(a.pool_cnt + q, static_cast<uint32_t>(cnt));
        float* ps = a.pool_scores + static_cast<size_t>(q) * a.pool_cap + slot0;
        int64_t* pi = a.pool_ids + static_cast<size_t>(q) * a.pool_cap + slot0;
        for (int i = 0; i < cnt; ++i) {
          const uint2 e = my[i];
          ps[i] = __uint_as_float(e.x);
          pi[i] = a.ids ? a.ids[e.y] : a.id_base + static_cast<int64_t>(e.y);
        }
      }
      if (a.stats) {
        for (int o = 16; o > 0; o >>= 1) {
          n_app += __shfl_xor_sync(kFull, n_app, o);
          n_cmp += __shfl_xor_sync(kFull, n_cmp, o);
          wait_ns = max(wait_ns, __shfl_xor_sync(kFull, wait_ns, o));
        }
        if (lane == 0) {
          atomicAdd(a.stats + 0, n_app);
          atomicAdd(a.stats + 1, n_cmp);
          atomicMax(a.stats + 2, wait_ns);
          atomicAdd(a.stats + 3, wait_ns);
        }
      }
    }
    __syncwarp();
    // all four epilogue warps done -> stop the service warp
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 128) bars->done = 1;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

cudaError_t mips_scan_prepare(uint32_t max_smem_bytes) {
  return cudaFuncSetAttribute(mips_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>(max_smem_bytes));
}

void launch_mips_scan(const CUtensorMap& tmap_q, const CUtensorMap& tmap_e, const ScanArgs& args,
                      int grid, uint32_t smem_bytes, cudaStream_t stream) {
  mips_scan_kernel<<<grid, kScanThreads, smem_bytes, stream>>>(tmap_q, tmap_e, args);
}

}  // namespace emdr2
