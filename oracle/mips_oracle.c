/*
 * oracle/mips_oracle.c — CPU restatement of the reference's brute-force MIPS search.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under emdr2_b200/ may import, link or execute this file; it is
 * the checker used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  The product path is the CUDA library and fails loudly without it.
 *
 * PARITY UNPINNED BY THE REFERENCE'S OWN TESTS: DevSinghSachan/emdr2 ships no tests, golden vectors
 * or fixtures.  This restatement is pinned instead against outputs of the
 * reference's own DistributedBruteForceIndex class executed in the build container (CPU tensors,
 * cuda device strings patched out): tests/golden/make_mips_golden.py -> tests/golden/mips_ref_*.npz,
 * checked by tests/test_oracle_golden.py.
 *
 * What it follows (reference @ edb8cf67):
 *   megatron/data/emdr2_index.py:268-305  DistributedBruteForceIndex.search_mips_index
 *       :281      C_i = Q · E_iᵀ                      (fp16 x fp16, fp32 accumulate in cuBLAS)
 *       :284-292  C[:, start:end] = C_i                (fp16 storage of the scores)
 *       :295      torch.topk(C, top_k, dim=1)          (sorted best first; tie order unspecified)
 *       :298-303  indices -> id_map[row]               (row -> doc id)
 *   megatron/data/emdr2_index.py:182-197  FaissMIPSIndex.search_mips_index -> IndexFlatIP.search:
 *       fp32 inner products of the stored rows, ids int64, sorted best first.
 *
 * Arithmetic of the oracle: score[q,i] = sum_j float(Q[q,j]) * float(E[i,j]) accumulated in double
 * and rounded once to fp32 ("exact" mode), optionally rounded again to fp16 to mimic :284
 * (round_fp16 != 0).  Ranking: (score descending, id ascending) — a deterministic refinement of the
 * reference's unspecified tie order.  tie_mask[q,r] = 1 when rank r's score is within
 * 2^-20*|score| of a neighbour or of the first excluded row (ids there are numerically ambiguous
 * for any fp32 implementation that accumulates in a different order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else { /* subnormal */
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while (!(man & 0x400u));
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp - 15 + 127) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

static float bf16_to_float(uint16_t b) {
  uint32_t bits = (uint32_t)b << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

/* round-to-nearest-even fp32 -> fp16 -> fp32 (mimics the fp16 score storage at emdr2_index.py:284) */
static float round_through_half(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  uint32_t sign = x & 0x80000000u;
  uint32_t ax = x & 0x7fffffffu;
  float out;
  if (ax >= 0x7f800000u) return f; /* inf / nan */
  if (ax >= 0x477ff000u) {         /* >= 65520 -> inf */
    uint32_t b = sign | 0x7f800000u;
    memcpy(&out, &b, 4);
    return out;
  }
  if (ax < 0x38800000u) { /* subnormal half: quantum 2^-24 */
    float a = fabsf(f);
    float q = nearbyintf(a * 16777216.0f) / 16777216.0f;
    return sign ? -q : q;
  }
  uint32_t lsb = (ax >> 13) & 1u;
  ax += 0x0fffu + lsb;
  ax &= ~0x1fffu;
  ax |= sign;
  memcpy(&out, &ax, 4);
  return out;
}

typedef struct {
  float s;
  int64_t id;
} hit_t;

/* a ranks before b under (score desc, id asc) */
static int before(float sa, int64_t ia, float sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}

/*
 * E: [n, d] row-major 16-bit, Q: [nq, d]; dtype 0 = fp16, 1 = bf16.
 * ids: [n] or NULL (id = id_base + row).
 * out_scores/out_ids: [nq, k]; missing ranks (n < k) get -inf / -1.
 * tie_mask: [nq, k] or NULL.
 * Single-threaded; the Python wrapper fans query ranges out over host threads.
 */
int oracle_mips_topk(const uint16_t* E, const int64_t* ids, int64_t id_base, int64_t n, int d,
                     int dtype, const uint16_t* Q, int nq, int k, int round_fp16,
                     float* out_scores, int64_t* out_ids, uint8_t* tie_mask) {
  if (n < 0 || d <= 0 || nq < 0 || k <= 0) return -1;
  if (nq == 0) return 0;
  double* qf = (double*)malloc(sizeof(double) * (size_t)d * (size_t)nq);
  double* rowf = (double*)malloc(sizeof(double) * (size_t)d);
  hit_t* best = (hit_t*)malloc(sizeof(hit_t) * (size_t)k * (size_t)nq);
  int* cnt = (int*)calloc((size_t)nq, sizeof(int));
  float* excl_s = (float*)malloc(sizeof(float) * (size_t)nq); /* best score among excluded rows */
  int* have_excl = (int*)calloc((size_t)nq, sizeof(int));
  if (!qf || !rowf || !best || !cnt || !excl_s || !have_excl) {
    free(qf); free(rowf); free(best); free(cnt); free(excl_s); free(have_excl);
    return -2;
  }
  for (size_t t = 0; t < (size_t)nq * d; ++t) qf[t] = dtype ? bf16_to_float(Q[t]) : half_to_float(Q[t]);
  for (int64_t i = 0; i < n; ++i) {
    const uint16_t* row = E + (size_t)i * d;
    if (dtype)
      for (int j = 0; j < d; ++j) rowf[j] = (double)bf16_to_float(row[j]);
    else
      for (int j = 0; j < d; ++j) rowf[j] = (double)half_to_float(row[j]);
    const int64_t id = ids ? ids[i] : id_base + i;
    for (int q = 0; q < nq; ++q) {
      const double* qv = qf + (size_t)q * d;
      double acc = 0.0;
      for (int j = 0; j < d; ++j) acc += qv[j] * rowf[j];
      float s = (float)acc;
      if (round_fp16) s = round_through_half(s);
      if (s != s) continue; /* NaN rows are never returned */
      hit_t* b = best + (size_t)q * k;
      int c = cnt[q];
      if (c == k && !before(s, id, b[k - 1].s, b[k - 1].id)) {
        if (!have_excl[q] || s > excl_s[q]) { excl_s[q] = s; have_excl[q] = 1; }
        continue;
      }
      if (c == k && (!have_excl[q] || b[k - 1].s > excl_s[q])) { excl_s[q] = b[k - 1].s; have_excl[q] = 1; }
      int pos = c < k ? c : k - 1;
      while (pos > 0 && before(s, id, b[pos - 1].s, b[pos - 1].id)) {
        b[pos] = b[pos - 1];
        --pos;
      }
      b[pos].s = s;
      b[pos].id = id;
      if (c < k) cnt[q] = c + 1;
    }
  }
  for (int q = 0; q < nq; ++q) {
    const hit_t* b = best + (size_t)q * k;
    const int c = cnt[q];
    for (int r = 0; r < k; ++r) {
      out_scores[(size_t)q * k + r] = r < c ? b[r].s : -INFINITY;
      out_ids[(size_t)q * k + r] = r < c ? b[r].id : -1;
      if (tie_mask) {
        uint8_t t = 0;
        if (r < c) {
          const float s = b[r].s;
          const float tol = ldexpf(fabsf(s), -20);
          if (r > 0 && fabsf(b[r - 1].s - s) <= tol) t = 1;
          if (r + 1 < c && fabsf(b[r + 1].s - s) <= tol) t = 1;
          if (have_excl[q] && fabsf(excl_s[q] - s) <= tol) t = 1;
        }
        tie_mask[(size_t)q * k + r] = t;
      }
    }
  }
  free(qf); free(rowf); free(best); free(cnt); free(excl_s); free(have_excl);
  return 0;
}

/* k-way merge restatement: [parts, nq, k] lists -> [nq, k] under (score desc, id asc); id < 0 = padding. */
int oracle_mips_merge(const float* scores, const int64_t* ids, int parts, int nq, int k,
                      float* out_scores, int64_t* out_ids) {
  for (int q = 0; q < nq; ++q) {
    int cnt = 0;
    for (int p = 0; p < parts; ++p)
      for (int j = 0; j < k; ++j) {
        size_t src = ((size_t)p * nq + q) * k + j;
        float s = scores[src];
        int64_t id = ids[src];
        if (id < 0 || s != s) continue;
        float* os = out_scores + (size_t)q * k;
        int64_t* oi = out_ids + (size_t)q * k;
        if (cnt == k && !before(s, id, os[k - 1], oi[k - 1])) continue;
        int pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && before(s, id, os[pos - 1], oi[pos - 1])) {
          os[pos] = os[pos - 1];
          oi[pos] = oi[pos - 1];
          --pos;
        }
        os[pos] = s;
        oi[pos] = id;
        if (cnt < k) ++cnt;
      }
    for (int r = cnt; r < k; ++r) {
      out_scores[(size_t)q * k + r] = -INFINITY;
      out_ids[(size_t)q * k + r] = -1;
    }
  }
  return 0;
}
