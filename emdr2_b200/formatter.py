"""Passage -> model-input formatting of the retrieve-and-read step (host side, integer work).

Mirrors reference megatron/model/emdr2_model.py:250-376 (`postprocess`,
`query_extended_context_t5_format`, `query_single_context_t5_format`) and
megatron/data/orqa_wiki_dataset.py:86-120 (`build_tokens_types_paddings_from_ids`, imported there
as `context_bert_format`): same names, argument order and outputs — element for element — but
written over preallocated numpy rows instead of growing Python lists, and returning one pinned
host block per tensor so the four uploads are single async copies.

Layouts produced (all int64):
  context ids / types   [B, K, S_ret]   [CLS] title [SEP] passage [SEP] pad...        (types all 0)
  extended              [B*K, S]        query title [SEP] passage(+neighbour fill) [SEP] pad...
  single                [B*K, S]        query title [SEP] passage [SEP] pad...
"""
import numpy as np
import torch


def build_tokens_types_paddings_from_ids(text_ids, max_seq_length, cls_id, sep_id, pad_id):
    """[CLS] text (cut to max_seq_length-2 tokens) [SEP] pad...; returns (ids, types, pad_mask)."""
    body = list(text_ids)[:max(0, max_seq_length - 2)]
    n = len(body) + 2
    ids = [cls_id] + body + [sep_id] + [pad_id] * (max_seq_length - n)
    types = [0] * n + [pad_id] * (max_seq_length - n)
    pad_mask = np.zeros(max(max_seq_length, n), dtype=np.int64)
    pad_mask[:n] = 1
    return ids, types, pad_mask


context_bert_format = build_tokens_types_paddings_from_ids


def _fill_context(context_doc_list, main_doc_idx, room):
    """The passage plus as much of its neighbours as fits in `room` tokens (emdr2_model.py:309-350)."""
    main = list(context_doc_list[main_doc_idx])
    if len(main) > room or len(context_doc_list) == 1:
        return main[:room]
    spare = room - len(main)
    if main_doc_idx == 0:                       # neighbours follow the passage
        tail = [t for doc in context_doc_list[1:] for t in doc]
        return main + tail[:spare]
    if main_doc_idx == -1:                      # neighbours precede the passage
        head = [t for doc in context_doc_list[:-1] for t in doc]
        if len(head) > spare:
            head = head[len(head) - spare + 1:]          # the reference keeps spare-1 tokens here
        return head + main
    left = list(context_doc_list[0])            # passage in the middle
    if len(left) > spare:
        return left[len(left) - spare + 1:] + main
    out = left + main
    if len(context_doc_list) == 3:
        out = out + list(context_doc_list[2])[:spare - len(left)]
    return out


def query_extended_context_t5_format(query_ids, title_ids, context_doc_list, main_doc_idx,
                                     max_seq_length, sep_id, pad_id):
    head = list(query_ids) + list(title_ids) + [sep_id]
    room = max(0, max_seq_length - len(head) - 1)
    enc = head + _fill_context(context_doc_list, main_doc_idx, room) + [sep_id]
    return enc + [pad_id] * (max_seq_length - len(enc))


def query_single_context_t5_format(query_ids, title_ids, context_ids, max_seq_length, sep_id, pad_id):
    enc = (list(query_ids) + list(title_ids) + [sep_id] + list(context_ids))[:max_seq_length - 1]
    enc.append(sep_id)
    return enc + [pad_id] * (max_seq_length - len(enc))


def _context_pieces(docs, main_doc_idx, room):
    """Array form of `_fill_context`: the list of int64 array slices whose concatenation is the
    passage plus as much of its neighbours as fits in `room` tokens."""
    main = docs[main_doc_idx]
    if len(main) > room or len(docs) == 1:
        return [main[:room]]
    spare = room - len(main)
    if main_doc_idx == 0:
        out = [main]
        for doc in docs[1:]:
            if spare <= 0:
                break
            out.append(doc[:spare])
            spare -= len(doc)
        return out
    if main_doc_idx == -1:
        head = docs[:-1]
        total = sum(len(d) for d in head)
        if total > spare:                     # the reference keeps the last spare-1 tokens here
            drop = total - spare + 1
            kept = []
            for d in head:
                if drop >= len(d):
                    drop -= len(d)
                    continue
                kept.append(d[drop:])
                drop = 0
            head = kept
        return list(head) + [main]
    left = docs[0]
    if len(left) > spare:
        return [left[len(left) - spare + 1:], main]
    out = [left, main]
    if len(docs) == 3:
        out.append(docs[2][:spare - len(left)])
    return out


def _put(row, pieces, limit):
    """Write the concatenation of `pieces` into row[:limit] (truncating); returns the length written."""
    pos = 0
    for p in pieces:
        n = min(len(p), limit - pos)
        if n <= 0:
            if pos >= limit:
                break
            continue                      # an empty piece (e.g. a neighbour cut to nothing)
        row[pos:pos + n] = p[:n]
        pos += n
    return pos


def postprocess_arrays(query_uid, query_ids_t5, query_ids_t5_len, topk_evidence_data, topk_retrievals,
                       seq_length_ret, seq_length, cls_id, sep_id, pad_id):
    """numpy version of `postprocess`: returns (context_ids [B,K,S_ret], context_types [B,K,S_ret],
    extended [B*K,S], single [B*K,S]) as int64 arrays.  Rows are assembled by slice assignment of
    int64 array pieces (passages may be given as lists or as arrays, e.g. straight from the
    memory-mapped token store); element-identical to the three `*_format` functions above."""
    uids = [int(u) for u in query_uid]
    bsz = len(uids)
    k_keep = int(topk_retrievals)
    ctx_ids = np.full((bsz, k_keep, seq_length_ret), pad_id, dtype=np.int64)
    ctx_types = np.zeros((bsz, k_keep, seq_length_ret), dtype=np.int64)
    extended = np.full((bsz * k_keep, seq_length), pad_id, dtype=np.int64)
    single = np.full((bsz * k_keep, seq_length), pad_id, dtype=np.int64)
    sep = np.array([sep_id], dtype=np.int64)
    cls = np.array([cls_id], dtype=np.int64)
    arr = lambda x: x if isinstance(x, np.ndarray) else np.asarray(x, dtype=np.int64)   # noqa: E731
    row = 0
    for bi, (qid, (topkids, text_list)) in enumerate(zip(uids, topk_evidence_data)):
        query = arr(query_ids_t5[bi])[:int(query_ids_t5_len[bi])]
        kept = 0
        for eid, (doc_list, main_idx, title_ids) in zip(topkids, text_list):
            if qid == eid or kept >= k_keep:        # drop the passage the question came from
                continue
            docs = [arr(d) for d in doc_list]
            title = arr(title_ids)
            passage = docs[main_idx]
            # [CLS] title [SEP] passage (cut to S_ret - 1) [SEP]
            n = _put(ctx_ids[bi, kept], (cls, title, sep, passage), seq_length_ret - 1)
            ctx_ids[bi, kept, n] = sep_id
            # query title [SEP] passage(+neighbours) [SEP]
            head_len = len(query) + len(title) + 1
            room = max(0, seq_length - head_len - 1)
            pieces = [query, title, sep] + _context_pieces(docs, main_idx, room) + [sep]
            if head_len + 1 > seq_length or sum(len(p) for p in pieces) > seq_length:
                raise ValueError("question + title do not fit in seq_length=%d" % seq_length)
            _put(extended[row], pieces, seq_length)
            # query title [SEP] passage, cut to S - 1, [SEP]
            n = _put(single[row], (query, title, sep, passage), seq_length - 1)
            single[row, n] = sep_id
            kept += 1
            row += 1
        if kept != k_keep:
            raise ValueError("query %d kept %d of %d passages (the reference would build a ragged "
                             "tensor here)" % (bi, kept, k_keep))
    return ctx_ids, ctx_types, extended, single


class FormatLengths(tuple):
    """(len_ctx, len_ext, len_one): the longest non-padding prefix of the context, extended and
    single-context layouts; `.rows` holds the same per row (three int32 host arrays) — known on the
    host for free, they let the towers trim and length-bucket their batches without a device sync."""

    def __new__(cls, max_len, row_len):
        self = super().__new__(cls, (int(max_len[0]), int(max_len[1]), int(max_len[2])))
        self.rows = (row_len[0], row_len[1], row_len[2])
        return self


class PackedTopk(object):
    """Retrieved evidence of a batch in the flat form `emdr2_format_passages_flat` reads: no token has
    been touched yet.  cand_begin int32 [B + 1], cand_id int64 [n], cand_meta int32 [n, 6] (title_len,
    n_docs, main_idx, doc_len x 3), piece_offset int64 [n, 4] (title | passages) into the two stores."""

    def __init__(self, cand_begin, cand_id, cand_meta, piece_offset, title_store, passage_store):
        if title_store.token_bytes != passage_store.token_bytes:
            raise TypeError("title and passage stores must share one token width")
        self.cand_begin = np.ascontiguousarray(cand_begin, dtype=np.int32)
        self.cand_id = np.ascontiguousarray(cand_id, dtype=np.int64)
        self.cand_meta = np.ascontiguousarray(cand_meta, dtype=np.int32).reshape(-1, 6)
        self.piece_offset = np.ascontiguousarray(piece_offset, dtype=np.int64).reshape(-1, 4)
        self.title_store, self.passage_store = title_store, passage_store

    def __len__(self):
        return self.cand_begin.shape[0] - 1

    def to_nested(self):
        """The reference's nested form (emdr2_model.py:457-468): [(ids, [(doc_list, main_idx, title)])]."""
        out = []
        for b in range(len(self)):
            ids, texts = [], []
            for c in range(self.cand_begin[b], self.cand_begin[b + 1]):
                tl, nd, main = (int(x) for x in self.cand_meta[c, :3])
                off = self.piece_offset[c]
                title = np.asarray(self.title_store.tokens[off[0]:off[0] + tl], dtype=np.int64)
                docs = [np.asarray(self.passage_store.tokens[off[1 + i]:off[1 + i] + int(self.cand_meta[c, 3 + i])],
                                   dtype=np.int64) for i in range(nd)]
                ids.append(int(self.cand_id[c]))
                texts.append((docs, main, title))
            out.append((ids, texts))
        return out


class _Staging(object):
    """Two alternating host blocks per output shape (pinned when the target is a CUDA device), each
    guarded by an event recorded after its upload, so a block is never rewritten while the copy engine
    may still be reading it."""

    def __init__(self):
        self.slots = {}

    def acquire(self, n_elems, pinned):
        key = (n_elems, pinned)
        ring = self.slots.setdefault(key, {"next": 0, "blocks": [None, None]})
        i = ring["next"]
        ring["next"] = 1 - i
        slot = ring["blocks"][i]
        if slot is None:
            slot = {"host": torch.empty(n_elems, dtype=torch.int64, pin_memory=pinned), "event": None}
            ring["blocks"][i] = slot
        if slot["event"] is not None:
            slot["event"].synchronize()
        return slot


_STAGING = _Staging()


def flatten_topk(topk_evidence_data):
    """Nested retriever output -> the flat arrays emdr2_format_passages takes: candidate ranges per
    question, candidate ids, [title_len, n_docs, main_idx, doc_len*3] per candidate, and one token
    buffer holding every candidate's title followed by its passages."""
    cand_begin = np.zeros(len(topk_evidence_data) + 1, dtype=np.int32)
    ids, meta, pieces = [], [], []
    for bi, (topkids, text_list) in enumerate(topk_evidence_data):
        for eid, (doc_list, main_idx, title_ids) in zip(topkids, text_list):
            nd = len(doc_list)
            if not 1 <= nd <= 3:
                raise ValueError("a passage comes with 1..3 paragraphs (itself and its neighbours), got %d" % nd)
            ids.append(eid)
            pieces.append(title_ids)
            pieces.extend(doc_list)
            meta.append((len(title_ids), nd, main_idx, len(doc_list[0]),
                         len(doc_list[1]) if nd > 1 else 0, len(doc_list[2]) if nd > 2 else 0))
        cand_begin[bi + 1] = len(ids)
    cand_id = np.asarray(ids, dtype=np.int64)
    cand_meta = np.asarray(meta, dtype=np.int32).reshape(-1, 6)
    if pieces:
        tokens = np.concatenate([p if isinstance(p, np.ndarray) else np.asarray(p, dtype=np.int64) for p in pieces])
        tokens = np.ascontiguousarray(tokens, dtype=np.int64)
    else:
        tokens = np.zeros(0, dtype=np.int64)
    return cand_begin, cand_id, cand_meta, tokens


def format_passages_native(query_uid, query_ids_t5, query_ids_t5_len, topk_evidence_data, topk_retrievals,
                           seq_length_ret, seq_length, cls_id, sep_id, pad_id, out=None):
    """`postprocess_arrays` through the library's host entry point emdr2_format_passages (one C call
    instead of B*K Python iterations).  `out`: optional flat int64 host buffer of
    rows*(2*seq_length_ret + 2*seq_length) elements to write into (e.g. pinned staging memory).
    Returns ((ctx_ids, ctx_types, extended, single) as numpy views of the buffer, FormatLengths).
    Raises ValueError where `postprocess_arrays` does."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    uid = np.ascontiguousarray(np.asarray(query_uid, dtype=np.int64).reshape(-1))
    bsz = int(uid.shape[0])
    k_keep = int(topk_retrievals)
    q = np.ascontiguousarray(np.asarray(query_ids_t5, dtype=np.int64).reshape(bsz, -1))
    q_len = np.ascontiguousarray(np.asarray(query_ids_t5_len, dtype=np.int64).reshape(-1))
    if len(topk_evidence_data) != bsz or q_len.shape[0] != bsz:
        raise ValueError("batch size mismatch between questions and retrieved evidence")
    packed = topk_evidence_data if isinstance(topk_evidence_data, PackedTopk) else None
    if packed is None:
        cand_begin, cand_id, cand_meta, tokens = flatten_topk(topk_evidence_data)
    rows = bsz * k_keep
    n_ret, n_seq = rows * seq_length_ret, rows * seq_length
    if out is None:
        out = np.empty(2 * n_ret + 2 * n_seq, dtype=np.int64)
    assert out.dtype == np.int64 and out.size >= 2 * n_ret + 2 * n_seq and out.flags["C_CONTIGUOUS"]
    ctx_ids = out[:n_ret].reshape(bsz, k_keep, seq_length_ret)
    extended = out[n_ret:n_ret + n_seq].reshape(rows, seq_length)
    single = out[n_ret + n_seq:n_ret + 2 * n_seq].reshape(rows, seq_length)
    ctx_types = out[n_ret + 2 * n_seq:2 * n_ret + 2 * n_seq].reshape(bsz, k_keep, seq_length_ret)
    max_len = np.zeros(3, dtype=np.int32)
    row_len = np.zeros((3, rows), dtype=np.int32)
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data)   # noqa: E731
    if packed is not None:
        # tokens are read in place from the flat stores (possibly memory-mapped files)
        tt, dt = packed.title_store.tokens, packed.passage_store.tokens
        rc = lib.emdr2_format_passages_flat(
            bsz, k_keep, ptr(uid), ptr(q), q.shape[1], ptr(q_len), ptr(packed.cand_begin), ptr(packed.cand_id),
            ptr(packed.cand_meta), ptr(packed.piece_offset), ptr(tt), tt.shape[0], ptr(dt), dt.shape[0],
            packed.passage_store.token_bytes, int(seq_length_ret), int(seq_length), int(cls_id), int(sep_id),
            int(pad_id), ptr(ctx_ids), ptr(ctx_types), ptr(extended), ptr(single), ptr(max_len), ptr(row_len))
    else:
        rc = lib.emdr2_format_passages(bsz, k_keep, ptr(uid), ptr(q), q.shape[1], ptr(q_len), ptr(cand_begin),
                                       ptr(cand_id), ptr(cand_meta), ptr(tokens), tokens.shape[0],
                                       int(seq_length_ret), int(seq_length), int(cls_id), int(sep_id), int(pad_id),
                                       ptr(ctx_ids), ptr(ctx_types), ptr(extended), ptr(single), ptr(max_len),
                                       ptr(row_len))
    if rc != 0:
        msg = lib.emdr2_last_error().decode("utf-8", "replace")
        raise ValueError(msg)
    return (ctx_ids, ctx_types, extended, single), FormatLengths(max_len, row_len)


def postprocess(query_uid, query_ids_t5, query_ids_t5_len, topk_evidence_data, topk_retrievals,
                seq_length_ret, seq_length, cls_id, sep_id, pad_id, device=None, return_lengths=False):
    """Drop-in for emdr2_model.py:250-303 with the tokenizer ids / lengths passed explicitly
    (the reference reads them from get_args()/get_t5_tokenizer()).  Returns four int64 tensors on
    `device` (default: current CUDA device).  The rows are assembled by the library's host entry
    point (`format_passages_native`) directly in pinned staging memory and uploaded with ONE
    asynchronous copy (the all-zero token-type tensor is created on the device)."""
    if torch.is_tensor(query_uid):
        query_uid = query_uid.tolist()
    if torch.is_tensor(query_ids_t5):
        query_ids_t5 = query_ids_t5.cpu().numpy()
    if torch.is_tensor(query_ids_t5_len):
        query_ids_t5_len = query_ids_t5_len.tolist()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    on_gpu = device.type == "cuda"
    bsz, k_keep = len(query_uid), int(topk_retrievals)
    rows = bsz * k_keep
    n_ret, n_seq = rows * seq_length_ret, rows * seq_length
    slot = _STAGING.acquire(2 * n_ret + 2 * n_seq, on_gpu)
    host = slot["host"]
    _, lengths = format_passages_native(query_uid, query_ids_t5, query_ids_t5_len, topk_evidence_data,
                                        topk_retrievals, seq_length_ret, seq_length, cls_id, sep_id, pad_id,
                                        out=host.numpy())
    if on_gpu:
        with torch.cuda.device(device):
            dev = host[:n_ret + 2 * n_seq].to(device, non_blocking=True)
            if slot["event"] is None:
                slot["event"] = torch.cuda.Event()
            slot["event"].record()
        types = torch.zeros((bsz, k_keep, seq_length_ret), dtype=torch.int64, device=device)
    else:
        dev = host[:n_ret + 2 * n_seq].clone()
        types = host[n_ret + 2 * n_seq:2 * n_ret + 2 * n_seq].clone().reshape(bsz, k_keep, seq_length_ret)
    out = (dev[:n_ret].reshape(bsz, k_keep, seq_length_ret), types,
           dev[n_ret:n_ret + n_seq].reshape(rows, seq_length),
           dev[n_ret + n_seq:n_ret + 2 * n_seq].reshape(rows, seq_length))
    if return_lengths:
        # longest non-padding prefix of each tensor, known on the host for free: lets the caller run
        # the towers on [:, :max_len] instead of the padded width without a device sync
        return out, lengths
    return out
