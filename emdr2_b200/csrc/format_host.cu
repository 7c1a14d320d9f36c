// Host-side passage -> model-input formatting (integer work, no device code).
//
// Replaces the per-(question, passage) Python loops of reference megatron/model/emdr2_model.py:250-303
// (`postprocess`) and the three list builders it calls:
//   context_bert_format              megatron/data/orqa_wiki_dataset.py:86-120  [CLS] title [SEP] passage [SEP]
//   query_extended_context_t5_format emdr2_model.py:306-361   query title [SEP] passage(+neighbours) [SEP]
//   query_single_context_t5_format   emdr2_model.py:364-376   query title [SEP] passage [SEP]
// B*K = 400 iterations of list surgery per rank per step there; one call over flat arrays here, written
// straight into the caller's (pinned) staging rows.  Element-identical to emdr2_b200/formatter.py's
// postprocess_arrays, which is itself pinned to the reference functions (tests/golden/formatter_ref.json).
#include <stdint.h>
#include <string.h>

#include "capi_common.cuh"

namespace {

using emdr2::capi::fail;

struct Span {
  const int64_t* p;
  int64_t n;
};

// Appends span s to row[pos..limit), truncating; returns the new position.
inline int64_t put(int64_t* row, int64_t pos, int64_t limit, Span s) {
  int64_t n = s.n < limit - pos ? s.n : limit - pos;
  if (n > 0) {
    memcpy(row + pos, s.p, static_cast<size_t>(n) * sizeof(int64_t));
    pos += n;
  }
  return pos;
}

inline Span head(Span s, int64_t n) {          // s[:n]
  if (n < 0) n = 0;
  return Span{s.p, n < s.n ? n : s.n};
}
inline Span tail_from(Span s, int64_t start) { // s[start:]
  if (start < 0) start = 0;
  if (start > s.n) start = s.n;
  return Span{s.p + start, s.n - start};
}

// The passage plus as much of its neighbours as fits in `room` tokens (emdr2_model.py:309-350);
// writes at most 4 spans into out, returns how many.
int context_pieces(const Span* docs, int n_docs, int main_idx, int64_t room, Span* out) {
  const Span main = docs[main_idx < 0 ? n_docs + main_idx : main_idx];
  if (main.n > room || n_docs == 1) {
    out[0] = head(main, room);
    return 1;
  }
  int64_t spare = room - main.n;
  int n = 0;
  if (main_idx == 0) {                          // neighbours follow the passage
    out[n++] = main;
    for (int i = 1; i < n_docs; ++i) {
      if (spare <= 0) break;
      out[n++] = head(docs[i], spare);
      spare -= docs[i].n;
    }
    return n;
  }
  if (main_idx == -1) {                         // neighbours precede the passage
    int64_t total = 0;
    for (int i = 0; i < n_docs - 1; ++i) total += docs[i].n;
    int64_t drop = total > spare ? total - spare + 1 : 0;   // the reference keeps spare-1 tokens here
    const bool cut = total > spare;
    for (int i = 0; i < n_docs - 1; ++i) {
      if (cut && drop >= docs[i].n) {
        drop -= docs[i].n;
        continue;
      }
      out[n++] = tail_from(docs[i], drop);
      drop = 0;
    }
    out[n++] = main;
    return n;
  }
  const Span left = docs[0];                    // passage in the middle
  if (left.n > spare) {
    out[n++] = tail_from(left, left.n - spare + 1);
    out[n++] = main;
    return n;
  }
  out[n++] = left;
  out[n++] = main;
  if (n_docs == 3) out[n++] = head(docs[2], spare - left.n);
  return n;
}

// Longest prefix of row[0..n) that does not end in pad_id (formatter.py's `longest`).
inline int32_t live_len(const int64_t* row, int64_t n, int64_t pad_id) {
  while (n > 0 && row[n - 1] == pad_id) --n;
  return static_cast<int32_t>(n);
}

}  // namespace

extern "C" int emdr2_format_passages(int32_t bsz, int32_t k_keep, const int64_t* query_uid,
                                     const int64_t* query_ids, int64_t query_stride,
                                     const int64_t* query_len, const int32_t* cand_begin,
                                     const int64_t* cand_id, const int32_t* cand_meta,
                                     const int64_t* tokens, int64_t n_tokens, int32_t seq_ret,
                                     int32_t seq, int64_t cls_id, int64_t sep_id, int64_t pad_id,
                                     int64_t* ctx_ids, int64_t* ctx_types, int64_t* extended,
                                     int64_t* single, int32_t* max_len, int32_t* row_len) {
  if (bsz < 0 || k_keep < 0 || seq_ret < 2 || seq < 2)
    return fail(EMDR2_EINVAL, "emdr2_format_passages: bad sizes (bsz %d, k %d, seq_ret %d, seq %d)",
                bsz, k_keep, seq_ret, seq);
  if (bsz > 0 && (!query_uid || !query_ids || !query_len || !cand_begin || !ctx_ids || !ctx_types ||
                  !extended || !single))
    return fail(EMDR2_EINVAL, "emdr2_format_passages: null pointer");
  const int64_t rows = static_cast<int64_t>(bsz) * k_keep;
  for (int64_t i = 0; i < rows * seq_ret; ++i) ctx_ids[i] = pad_id;
  memset(ctx_types, 0, static_cast<size_t>(rows) * seq_ret * sizeof(int64_t));
  for (int64_t i = 0; i < rows * seq; ++i) extended[i] = pad_id;
  for (int64_t i = 0; i < rows * seq; ++i) single[i] = pad_id;
  int32_t longest[3] = {0, 0, 0};

  // token offset of every candidate's first piece: candidates are laid out back to back as
  // title, doc 0, doc 1, doc 2 (absent docs have length 0)
  int64_t offset = 0;
  int64_t row = 0;
  const Span cls{&cls_id, 1}, sep{&sep_id, 1};
  for (int32_t b = 0; b < bsz; ++b) {
    if (query_len[b] < 0 || query_len[b] > query_stride)
      return fail(EMDR2_EINVAL, "emdr2_format_passages: query %d has length %lld outside [0, %lld]", b,
                  static_cast<long long>(query_len[b]), static_cast<long long>(query_stride));
    const Span query{query_ids + static_cast<int64_t>(b) * query_stride, query_len[b]};
    int32_t kept = 0;
    for (int32_t c = cand_begin[b]; c < cand_begin[b + 1]; ++c) {
      const int32_t* meta = cand_meta + static_cast<int64_t>(c) * 6;
      const int32_t title_len = meta[0], n_docs = meta[1], main_idx = meta[2];
      if (title_len < 0 || n_docs < 1 || n_docs > 3 || main_idx < -n_docs || main_idx >= n_docs ||
          meta[3] < 0 || meta[4] < 0 || meta[5] < 0)
        return fail(EMDR2_EINVAL, "emdr2_format_passages: candidate %d has malformed metadata", c);
      const Span title{tokens + offset, title_len};
      Span docs[3];
      int64_t o = offset + title_len;
      for (int i = 0; i < 3; ++i) {
        docs[i] = Span{tokens + o, i < n_docs ? meta[3 + i] : 0};
        o += meta[3 + i];
      }
      offset = o;
      if (offset > n_tokens)
        return fail(EMDR2_EINVAL, "emdr2_format_passages: token buffer too short (%lld > %lld)",
                    static_cast<long long>(offset), static_cast<long long>(n_tokens));
      if (cand_id[c] == query_uid[b] || kept >= k_keep) continue;   // drop the question's own passage
      const Span passage = docs[main_idx < 0 ? n_docs + main_idx : main_idx];

      // [CLS] title [SEP] passage (cut to S_ret - 1) [SEP]
      int64_t* r0 = ctx_ids + row * seq_ret;
      int64_t n = put(r0, 0, seq_ret - 1, cls);
      n = put(r0, n, seq_ret - 1, title);
      n = put(r0, n, seq_ret - 1, sep);
      n = put(r0, n, seq_ret - 1, passage);
      r0[n] = sep_id;
      const int32_t l0 = live_len(r0, n + 1, pad_id);
      if (l0 > longest[0]) longest[0] = l0;
      if (row_len) row_len[row] = l0;

      // query title [SEP] passage(+neighbours) [SEP]
      const int64_t head_len = query.n + title.n + 1;
      const int64_t room = seq - head_len - 1 > 0 ? seq - head_len - 1 : 0;
      Span ctx[4];
      const int n_ctx = context_pieces(docs, n_docs, main_idx, room, ctx);
      int64_t total = head_len + 1;
      for (int i = 0; i < n_ctx; ++i) total += ctx[i].n;
      if (head_len + 1 > seq || total > seq)
        return fail(EMDR2_EINVAL, "question + title do not fit in seq_length=%d", seq);
      int64_t* r1 = extended + row * seq;
      n = put(r1, 0, seq, query);
      n = put(r1, n, seq, title);
      n = put(r1, n, seq, sep);
      for (int i = 0; i < n_ctx; ++i) n = put(r1, n, seq, ctx[i]);
      n = put(r1, n, seq, sep);
      const int32_t l1 = live_len(r1, n, pad_id);
      if (l1 > longest[1]) longest[1] = l1;
      if (row_len) row_len[rows + row] = l1;

      // query title [SEP] passage, cut to S - 1, [SEP]
      int64_t* r2 = single + row * seq;
      n = put(r2, 0, seq - 1, query);
      n = put(r2, n, seq - 1, title);
      n = put(r2, n, seq - 1, sep);
      n = put(r2, n, seq - 1, passage);
      r2[n] = sep_id;
      const int32_t l2 = live_len(r2, n + 1, pad_id);
      if (l2 > longest[2]) longest[2] = l2;
      if (row_len) row_len[2 * rows + row] = l2;

      ++kept;
      ++row;
    }
    if (kept != k_keep)
      return fail(EMDR2_EINVAL,
                  "query %d kept %d of %d passages (the reference would build a ragged tensor here)", b,
                  kept, k_keep);
  }
  if (max_len) {
    max_len[0] = longest[0];
    max_len[1] = longest[1];
    max_len[2] = longest[2];
  }
  return EMDR2_OK;
}
