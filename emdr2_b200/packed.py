"""Token-packed (variable-length) execution of the towers: host-side bookkeeping.

The reference runs every tower on padded rectangles ([400, 256] contexts, [400, 512] reader inputs per
question batch, megatron/model/emdr2_model.py:118-120,148-149) and masks the padding inside attention.
The formatter knows every row's real length on the host, so the forward path can instead lay the
sequences back to back: activations are [T, h] with T = sum of the lengths — every GEMM, LayerNorm and
embedding row is a real token — and attention runs over an explicit work list in ONE launch per layer
(csrc/attention_varlen.cu).  This module builds what that needs from the lengths alone (numpy, then one
pinned upload): the gather index that packs the padded id matrix, the position ids, and the work items
of self- and cross-attention.  Results at real tokens equal the rectangular run's (a padding key has
probability exactly 0 there); padding positions simply do not exist.
"""
import numpy as np
import torch

TILE = 128


def _upload(arrays, device):
    """One pinned staging buffer, one asynchronous copy for a handful of int32 arrays."""
    sizes = [int(a.size) for a in arrays]
    pad = [(-n) % 4 for n in sizes]                      # keep every part 16-byte aligned
    total = sum(n + p for n, p in zip(sizes, pad))
    host = torch.empty(max(total, 4), dtype=torch.int32).pin_memory()
    view = host.numpy()
    offs, off = [], 0
    for a, n, p in zip(arrays, sizes, pad):
        view[off:off + n] = a.reshape(-1)
        offs.append(off)
        off += n + p
    dev = host.to(device, non_blocking=True)
    return [dev[o:o + n].view(a.shape) for a, o, n in zip(arrays, offs, sizes)], host


def self_attention_items(lens, cu, heads):
    """int32 [n, 8] work items (q_row0, q_valid, k_row0, k_len, head, o_row0, lse_idx0, 0) of the self-attention
    of every sequence: one per (sequence, 128-query tile, head), sorted by decreasing cost."""
    lens = np.asarray(lens, dtype=np.int64)
    tiles = -(-lens // TILE)
    seq = np.repeat(np.arange(lens.shape[0]), tiles)                       # sequence of every tile
    first = np.cumsum(tiles) - tiles
    t_in_seq = np.arange(int(tiles.sum())) - np.repeat(first, tiles)
    q_row0 = cu[seq] + t_in_seq * TILE
    q_valid = np.minimum(TILE, lens[seq] - t_in_seq * TILE)
    n_tiles = seq.shape[0]
    items = np.zeros((n_tiles, heads, 8), dtype=np.int64)
    items[:, :, 0] = q_row0[:, None]
    items[:, :, 1] = q_valid[:, None]
    items[:, :, 2] = cu[seq][:, None]
    items[:, :, 3] = lens[seq][:, None]
    items[:, :, 4] = np.arange(heads)[None, :]
    items[:, :, 5] = q_row0[:, None]
    items = items.reshape(-1, 8)
    order = np.argsort(-(items[:, 3] * 4 + items[:, 1] // 32), kind="stable")   # long key ranges first
    return items[order].astype(np.int32)


class PackedBatch(object):
    """b sequences of `lens` tokens packed into T rows (in their original order)."""

    def __init__(self, lens, padded_width, heads, device):
        lens = np.maximum(np.asarray(lens, dtype=np.int64).reshape(-1), 1)      # an empty row keeps its first slot
        self.lens = lens
        self.b = int(lens.shape[0])
        self.cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        self.T = int(self.cu[-1])
        self.heads = heads
        starts = np.arange(self.b, dtype=np.int64) * int(padded_width)
        within = np.arange(self.T, dtype=np.int64) - np.repeat(self.cu[:-1], lens)
        gather = np.repeat(starts, lens) + within                               # index into ids.view(-1)
        items = self_attention_items(lens, self.cu, heads)
        self.n_items = int(items.shape[0])
        (self.gather, self.pos_ids, self.items, self.first_rows), self._staging = _upload(
            [gather.astype(np.int32), within.astype(np.int32), items, self.cu[:-1].astype(np.int32)], device)
        self.device = device
        # algorithmic work of one self-attention pass: 4 * heads * 64 * sum(len^2)
        self.attention_flops = 4.0 * heads * 64 * float((lens.astype(np.float64) ** 2).sum())


class PackedStates(object):
    """Encoder states of a packed batch: `states` [T, h], sequence i = rows [cu[i], cu[i+1]).  What the FiD
    decoder attends over: question q's keys are the rows of its `group` consecutive sequences, contiguous."""

    def __init__(self, states, lens, cu, group=1):
        self.states, self.lens, self.cu = states, np.asarray(lens), np.asarray(cu)
        self.group = int(group)          # sequences per key set (FiD: the top-k passages of one question)
        self._cross = {}

    def index_select(self, dim, index):
        """Key sets `index` (with repetition), like Tensor.index_select(0, ...) on the padded [sets, keys, h] states:
        what the reference's beam search does to carry the encoder states of a question along with each of its
        hypotheses (search_strategy.py:91-103).  The cached decode loop never needs it."""
        if dim != 0:
            raise ValueError("packed encoder states are selected by key set (dim 0)")
        sets = np.asarray(index.tolist() if torch.is_tensor(index) else index, dtype=np.int64)
        g = self.group
        seqs = (sets[:, None] * g + np.arange(g)[None, :]).reshape(-1)
        lens = self.lens[seqs]
        cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        rows = np.repeat(self.cu[seqs] - cu[:-1], lens) + np.arange(int(cu[-1]))
        picked = self.states.index_select(0, torch.from_numpy(rows).to(self.states.device))
        return PackedStates(picked, lens, cu, group=g)

    @property
    def shape(self):
        return (len(self.lens), int(self.lens.max()), self.states.shape[1])

    def to_padded(self, width=None):
        """[b, width, h] rectangle with zeros at padding (the reference's layout), for callers that want it."""
        b, h = len(self.lens), self.states.shape[1]
        width = int(width or self.lens.max())
        out = self.states.new_zeros((b * width, h))
        idx = np.repeat(np.arange(b) * width - self.cu[:-1], self.lens) + np.arange(int(self.cu[-1]))
        out.index_copy_(0, torch.from_numpy(idx).to(self.states.device), self.states)
        return out.view(b, width, h)

    def cross_plan(self, group, sq, heads, target_items=888):
        """Work items of a cross-attention in which every `group` consecutive sequences form one key set and each
        key set is attended by `sq` query rows (FiD: group = top-k passages, sq = decoder length).  Key sets are
        cut into ranges of `chunk` keys that run as separate items; their partial outputs are merged by the lse
        weights (autograd.cross_attention_packed).  Cached per (group, sq, heads)."""
        key = (int(group), int(sq), int(heads))
        plan = self._cross.get(key)
        if plan is None:
            plan = CrossPlan(self, *key, target_items=target_items)
            self._cross[key] = plan
        return plan


class CrossPlan(object):
    def __init__(self, states, group, sq, heads, target_items=888):
        cu, dev = states.cu, states.states.device
        n_sets = (len(states.lens)) // group
        if n_sets * group != len(states.lens):
            raise ValueError("%d sequences do not split into key sets of %d" % (len(states.lens), group))
        k0 = cu[np.arange(n_sets) * group]
        k1 = cu[(np.arange(n_sets) + 1) * group]
        klen = k1 - k0
        q_tiles = -(-sq // TILE)
        blocks = int((-(-klen // TILE)).sum())
        per = max(2, -(-blocks * heads * q_tiles // target_items))          # key blocks per range
        chunk = per * TILE
        n_chunks = -(-klen // chunk)
        slot0 = np.cumsum(n_chunks) - n_chunks
        n_slots = int(n_chunks.sum())
        self.n_sets, self.sq, self.heads, self.n_slots, self.max_chunks = n_sets, sq, heads, n_slots, int(n_chunks.max())
        rows = []
        for s in range(n_sets):
            for c in range(int(n_chunks[s])):
                kr0 = int(k0[s] + c * chunk)
                kl = int(min(chunk, k1[s] - kr0))
                slot = int(slot0[s] + c)
                for t in range(q_tiles):
                    qv = min(TILE, sq - t * TILE)
                    for hd in range(heads):
                        rows.append((s * sq + t * TILE, qv, kr0, kl, hd, slot * sq + t * TILE,
                                     (slot * heads + hd) * sq + t * TILE, 0))
        items = np.asarray(rows, dtype=np.int64)
        order = np.argsort(-items[:, 3], kind="stable")
        items = items[order].astype(np.int32)
        self.n_items = int(items.shape[0])
        slot_index = np.full((n_sets, self.max_chunks), n_slots, dtype=np.int32)     # n_slots = the empty slot
        for s in range(n_sets):
            slot_index[s, :int(n_chunks[s])] = slot0[s] + np.arange(int(n_chunks[s]))
        (self.items, self.slot_index), self._staging = _upload([items, slot_index], dev)
        self.attention_flops = 4.0 * heads * 64 * sq * float(klen.sum())
