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

// A run of tokens in the caller's store (any integer width) that is copied, widened to int64.
template <typename T>
struct Span {
  const T* p;
  int64_t n;
};

// Appends span s to row[pos..limit), truncating; returns the new position.
template <typename T>
inline int64_t put(int64_t* row, int64_t pos, int64_t limit, Span<T> s) {
  const int64_t n = s.n < limit - pos ? s.n : limit - pos;
  if (n > 0) {
    if (sizeof(T) == sizeof(int64_t)) {
      memcpy(row + pos, s.p, static_cast<size_t>(n) * sizeof(int64_t));
    } else {
      for (int64_t i = 0; i < n; ++i) row[pos + i] = static_cast<int64_t>(s.p[i]);
    }
    pos += n;
  }
  return pos;
}
inline int64_t put_one(int64_t* row, int64_t pos, int64_t limit, int64_t v) {
  if (pos < limit) row[pos++] = v;
  return pos;
}

template <typename T>
inline Span<T> head(Span<T> s, int64_t n) {          // s[:n]
  if (n < 0) n = 0;
  return Span<T>{s.p, n < s.n ? n : s.n};
}
template <typename T>
inline Span<T> tail_from(Span<T> s, int64_t start) { // s[start:]
  if (start < 0) start = 0;
  if (start > s.n) start = s.n;
  return Span<T>{s.p + start, s.n - start};
}

// The passage plus as much of its neighbours as fits in `room` tokens (emdr2_model.py:309-350);
// writes at most 4 spans into out, returns how many.
template <typename T>
int context_pieces(const Span<T>* docs, int n_docs, int main_idx, int64_t room, Span<T>* out) {
  const Span<T> main = docs[main_idx < 0 ? n_docs + main_idx : main_idx];
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
  const Span<T> left = docs[0];                 // passage in the middle
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

struct FormatArgs {
  int32_t bsz, k_keep;
  const int64_t* query_uid;
  const int64_t* query_ids;
  int64_t query_stride;
  const int64_t* query_len;
  const int32_t* cand_begin;
  const int64_t* cand_id;
  const int32_t* cand_meta;      // [n_cand, 6]: title_len, n_docs, main_idx, doc_len[3]
  const int64_t* piece_offset;   // [n_cand, 4] element offsets (title | doc 0..2), or NULL = back to back
  int64_t n_title_tokens, n_doc_tokens;
  int32_t seq_ret, seq;
  int64_t cls_id, sep_id, pad_id;
  int64_t *ctx_ids, *ctx_types, *extended, *single;
  int32_t *max_len, *row_len;
};

template <typename T>
int format_impl(const FormatArgs& a, const T* title_tokens, const T* doc_tokens) {
  const int32_t bsz = a.bsz, k_keep = a.k_keep, seq_ret = a.seq_ret, seq = a.seq;
  const int64_t pad_id = a.pad_id, sep_id = a.sep_id, cls_id = a.cls_id;
  const int64_t rows = static_cast<int64_t>(bsz) * k_keep;
  for (int64_t i = 0; i < rows * seq_ret; ++i) a.ctx_ids[i] = pad_id;
  memset(a.ctx_types, 0, static_cast<size_t>(rows) * seq_ret * sizeof(int64_t));
  for (int64_t i = 0; i < rows * seq; ++i) a.extended[i] = pad_id;
  for (int64_t i = 0; i < rows * seq; ++i) a.single[i] = pad_id;
  int32_t longest[3] = {0, 0, 0};

  int64_t running = 0;   // back-to-back layout: title, doc 0, doc 1, doc 2 per candidate
  int64_t row = 0;
  for (int32_t b = 0; b < bsz; ++b) {
    if (a.query_len[b] < 0 || a.query_len[b] > a.query_stride)
      return fail(EMDR2_EINVAL, "emdr2_format_passages: query %d has length %lld outside [0, %lld]", b,
                  static_cast<long long>(a.query_len[b]), static_cast<long long>(a.query_stride));
    const Span<int64_t> query{a.query_ids + static_cast<int64_t>(b) * a.query_stride, a.query_len[b]};
    int32_t kept = 0;
    for (int32_t c = a.cand_begin[b]; c < a.cand_begin[b + 1]; ++c) {
      const int32_t* meta = a.cand_meta + static_cast<int64_t>(c) * 6;
      const int32_t title_len = meta[0], n_docs = meta[1], main_idx = meta[2];
      if (title_len < 0 || n_docs < 1 || n_docs > 3 || main_idx < -n_docs || main_idx >= n_docs ||
          meta[3] < 0 || meta[4] < 0 || meta[5] < 0)
        return fail(EMDR2_EINVAL, "emdr2_format_passages: candidate %d has malformed metadata", c);
      Span<T> title, docs[3];
      if (a.piece_offset) {
        const int64_t* off = a.piece_offset + static_cast<int64_t>(c) * 4;
        if (off[0] < 0 || off[0] + title_len > a.n_title_tokens)
          return fail(EMDR2_EINVAL, "emdr2_format_passages: title of candidate %d lies outside the title store", c);
        title = Span<T>{title_tokens + off[0], title_len};
        for (int i = 0; i < 3; ++i) {
          const int64_t len = i < n_docs ? meta[3 + i] : 0;
          if (len > 0 && (off[1 + i] < 0 || off[1 + i] + len > a.n_doc_tokens))
            return fail(EMDR2_EINVAL, "emdr2_format_passages: passage %d of candidate %d lies outside the store", i, c);
          docs[i] = Span<T>{doc_tokens + (len > 0 ? off[1 + i] : 0), len};
        }
      } else {
        title = Span<T>{doc_tokens + running, title_len};
        int64_t o = running + title_len;
        for (int i = 0; i < 3; ++i) {
          docs[i] = Span<T>{doc_tokens + o, i < n_docs ? meta[3 + i] : 0};
          o += meta[3 + i];
        }
        running = o;
        if (running > a.n_doc_tokens)
          return fail(EMDR2_EINVAL, "emdr2_format_passages: token buffer too short (%lld > %lld)",
                      static_cast<long long>(running), static_cast<long long>(a.n_doc_tokens));
      }
      if (a.cand_id[c] == a.query_uid[b] || kept >= k_keep) continue;   // drop the question's own passage
      const Span<T> passage = docs[main_idx < 0 ? n_docs + main_idx : main_idx];

      // [CLS] title [SEP] passage (cut to S_ret - 1) [SEP]
      int64_t* r0 = a.ctx_ids + row * seq_ret;
      int64_t n = put_one(r0, 0, seq_ret - 1, cls_id);
      n = put(r0, n, seq_ret - 1, title);
      n = put_one(r0, n, seq_ret - 1, sep_id);
      n = put(r0, n, seq_ret - 1, passage);
      r0[n] = sep_id;
      const int32_t l0 = live_len(r0, n + 1, pad_id);
      if (l0 > longest[0]) longest[0] = l0;
      if (a.row_len) a.row_len[row] = l0;

      // query title [SEP] passage(+neighbours) [SEP]
      const int64_t head_len = query.n + title.n + 1;
      const int64_t room = seq - head_len - 1 > 0 ? seq - head_len - 1 : 0;
      Span<T> ctx[4];
      const int n_ctx = context_pieces(docs, n_docs, main_idx, room, ctx);
      int64_t total = head_len + 1;
      for (int i = 0; i < n_ctx; ++i) total += ctx[i].n;
      if (head_len + 1 > seq || total > seq)
        return fail(EMDR2_EINVAL, "question + title do not fit in seq_length=%d", seq);
      int64_t* r1 = a.extended + row * seq;
      n = put(r1, 0, seq, query);
      n = put(r1, n, seq, title);
      n = put_one(r1, n, seq, sep_id);
      for (int i = 0; i < n_ctx; ++i) n = put(r1, n, seq, ctx[i]);
      n = put_one(r1, n, seq, sep_id);
      const int32_t l1 = live_len(r1, n, pad_id);
      if (l1 > longest[1]) longest[1] = l1;
      if (a.row_len) a.row_len[rows + row] = l1;

      // query title [SEP] passage, cut to S - 1, [SEP]
      int64_t* r2 = a.single + row * seq;
      n = put(r2, 0, seq - 1, query);
      n = put(r2, n, seq - 1, title);
      n = put_one(r2, n, seq - 1, sep_id);
      n = put(r2, n, seq - 1, passage);
      r2[n] = sep_id;
      const int32_t l2 = live_len(r2, n + 1, pad_id);
      if (l2 > longest[2]) longest[2] = l2;
      if (a.row_len) a.row_len[2 * rows + row] = l2;

      ++kept;
      ++row;
    }
    if (kept != k_keep)
      return fail(EMDR2_EINVAL,
                  "query %d kept %d of %d passages (the reference would build a ragged tensor here)", b,
                  kept, k_keep);
  }
  if (a.max_len) {
    a.max_len[0] = longest[0];
    a.max_len[1] = longest[1];
    a.max_len[2] = longest[2];
  }
  return EMDR2_OK;
}

int check_common(const FormatArgs& a) {
  if (a.bsz < 0 || a.k_keep < 0 || a.seq_ret < 2 || a.seq < 2)
    return fail(EMDR2_EINVAL, "emdr2_format_passages: bad sizes (bsz %d, k %d, seq_ret %d, seq %d)", a.bsz,
                a.k_keep, a.seq_ret, a.seq);
  if (a.bsz > 0 && (!a.query_uid || !a.query_ids || !a.query_len || !a.cand_begin || !a.ctx_ids ||
                    !a.ctx_types || !a.extended || !a.single))
    return fail(EMDR2_EINVAL, "emdr2_format_passages: null pointer");
  return EMDR2_OK;
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
  const FormatArgs a{bsz, k_keep, query_uid, query_ids, query_stride, query_len, cand_begin, cand_id, cand_meta,
                     nullptr, n_tokens, n_tokens, seq_ret, seq, cls_id, sep_id, pad_id, ctx_ids, ctx_types,
                     extended, single, max_len, row_len};
  const int rc = check_common(a);
  if (rc != EMDR2_OK) return rc;
  return format_impl<int64_t>(a, tokens, tokens);
}

extern "C" int emdr2_format_passages_flat(int32_t bsz, int32_t k_keep, const int64_t* query_uid,
                                          const int64_t* query_ids, int64_t query_stride,
                                          const int64_t* query_len, const int32_t* cand_begin,
                                          const int64_t* cand_id, const int32_t* cand_meta,
                                          const int64_t* piece_offset, const void* title_tokens,
                                          int64_t n_title_tokens, const void* doc_tokens,
                                          int64_t n_doc_tokens, int32_t token_bytes, int32_t seq_ret,
                                          int32_t seq, int64_t cls_id, int64_t sep_id, int64_t pad_id,
                                          int64_t* ctx_ids, int64_t* ctx_types, int64_t* extended,
                                          int64_t* single, int32_t* max_len, int32_t* row_len) {
  const FormatArgs a{bsz, k_keep, query_uid, query_ids, query_stride, query_len, cand_begin, cand_id, cand_meta,
                     piece_offset, n_title_tokens, n_doc_tokens, seq_ret, seq, cls_id, sep_id, pad_id, ctx_ids,
                     ctx_types, extended, single, max_len, row_len};
  const int rc = check_common(a);
  if (rc != EMDR2_OK) return rc;
  if (!piece_offset || (cand_begin && bsz > 0 && cand_begin[bsz] > 0 && (!title_tokens || !doc_tokens)))
    return fail(EMDR2_EINVAL, "emdr2_format_passages_flat: null store / offset pointer");
  switch (token_bytes) {
    case 2:
      return format_impl<uint16_t>(a, static_cast<const uint16_t*>(title_tokens),
                                   static_cast<const uint16_t*>(doc_tokens));
    case 4:
      return format_impl<int32_t>(a, static_cast<const int32_t*>(title_tokens),
                                  static_cast<const int32_t*>(doc_tokens));
    case 8:
      return format_impl<int64_t>(a, static_cast<const int64_t*>(title_tokens),
                                  static_cast<const int64_t*>(doc_tokens));
    default:
      return fail(EMDR2_EINVAL, "token_bytes must be 2 (uint16), 4 (int32) or 8 (int64), got %d", token_bytes);
  }
}
