"""Evidence-embedding store: dense blocks in memory, a flat shard-addressable file format, and the
reference's pickle as an import/export format.

Stands where ``OpenRetreivalDataStore`` does (reference megatron/data/emdr2_index.py:16-100) with the
same method names and error behaviour (``add_block_data`` raises ValueError on an overwrite, :56-61;
``merge_shards_and_save`` asserts that shards do not overlap, :88-90), but is built differently:

* In memory the rows are BLOCKS — ``[n_i]`` int64 id arrays and ``[n_i, d]`` float16 matrices in
  insertion order — not a ``{doc_id: np.float16[d]}`` dict.  ``add_block_data`` appends one block
  (no per-row Python insert); ``to_arrays`` concatenates (no 21 M-iteration loop, which is where
  the reference's ``add_embed_data`` spends minutes, :245-260).  ``embed_data`` remains available as
  a property that materialises the reference's dict on demand, for callers that index it.
* On disk the primary format is flat: ``<path>.rows.f16`` = 16-byte header (N, d as int64) + ``[N, d]``
  float16 row-major, ``<path>.ids.i64`` = ``[N]`` int64.  ``load_from_file(row_range=(lo, hi))`` memory-maps
  it and touches only that slice, which is how each rank of ``B200BruteForceIndex.update_index`` loads
  its own ``torch.chunk`` range instead of unpickling 32 GB.
* The reference's pickle (``{'embed_data': {int: np.float16[d]}}``, :33-36) is read when no flat file is
  present and written by ``save_shard`` / ``merge_shards_and_save`` with ``format='pickle'`` (the default,
  so that a reference trainer or indexer on the other side of the hand-over keeps working);
  ``format='flat'`` makes both write the flat layout instead.
"""
import os
import pickle
import shutil

import numpy as np


class EvidenceStore(object):
    """Serializable holder of evidence embeddings keyed by doc id (reference: OpenRetreivalDataStore)."""

    def __init__(self, embedding_path=None, load_from_path=True, rank=None, format="pickle"):
        if embedding_path is None:
            raise ValueError("embedding_path is required (the reference reads args.embedding_path here)")
        if format not in ("pickle", "flat"):
            raise ValueError("format must be 'pickle' or 'flat'")
        self.embedding_path = embedding_path
        self.rank = 0 if rank is None else rank
        self.format = format
        self._id_blocks, self._row_blocks = [], []
        self._seen = _IdSet()
        self.loaded_range = None            # (lo, hi, N) when only a slice of the file was loaded
        block_data_name = os.path.splitext(self.embedding_path)[0]
        self.temp_dir_name = block_data_name + '_tmp'
        if load_from_path:
            self.load_from_file()

    # ------------------------------------------------------------------ reference surface
    def __len__(self):
        return sum(int(b.shape[0]) for b in self._id_blocks)

    @property
    def embed_data(self):
        """The reference's ``{doc_id: float16[d]}`` view, built on demand (rows are views, not copies)."""
        ids, rows = self.to_arrays()
        return dict(zip(ids.tolist(), rows))

    @embed_data.setter
    def embed_data(self, mapping):
        self.clear()
        if mapping:
            ids, rows = dict_to_arrays(mapping)
            self._append(ids, rows, allow_overwrite=True)

    def state(self):
        return {'embed_data': self.embed_data}

    def clear(self):
        """Drop the embeddings (the index owns a copy once add_embed_data has run; :38-43)."""
        self._id_blocks, self._row_blocks = [], []
        self._seen = _IdSet()
        self.loaded_range = None

    def load_from_file(self, row_range=None):
        """Populate from disk (:45-54): the flat files when present (memory-mapped; ``row_range`` loads
        one slice), else the reference's pickle."""
        self.clear()
        if flat_exists(self.embedding_path):
            ids, rows = load_flat(self.embedding_path, row_range)
            n_total = flat_shape(self.embedding_path)[0]
            self._id_blocks, self._row_blocks = [ids], [rows]
            self._seen = None               # a file written by this class holds no duplicates; checked lazily
            if row_range is not None:
                self.loaded_range = (int(row_range[0]), int(row_range[0]) + int(ids.shape[0]), n_total)
            return
        with open(self.embedding_path, 'rb') as f:
            state_dict = pickle.load(f)
        ids, rows = dict_to_arrays(state_dict['embed_data'])
        if row_range is not None:
            lo, hi = row_range
            self.loaded_range = (int(lo), min(int(hi), int(ids.shape[0])), int(ids.shape[0]))
            ids, rows = ids[lo:hi], rows[lo:hi]
        self._append(ids, rows, allow_overwrite=True)

    def add_block_data(self, row_id, block_embeds, allow_overwrite=False):
        """Append rows as float16 keyed by doc id (:56-61); one vectorised block, not a per-row insert."""
        ids = np.asarray(list(row_id) if not hasattr(row_id, "__array__") else row_id, dtype=np.int64).reshape(-1)
        rows = np.asarray(block_embeds)
        if rows.dtype != np.float16:
            rows = rows.astype(np.float16)
        if rows.ndim != 2 or rows.shape[0] != ids.shape[0]:
            raise ValueError("block_embeds must be [len(row_id), d]")
        self._append(ids, np.ascontiguousarray(rows), allow_overwrite)

    def _append(self, ids, rows, allow_overwrite):
        if ids.shape[0] == 0:
            return
        if self._seen is None:              # loaded from a flat file: index what is there first
            self._seen = _IdSet()
            for b in self._id_blocks:
                self._seen.add(np.asarray(b))
        if self._seen.any_known(ids) or np.unique(ids).shape[0] != ids.shape[0]:
            if not allow_overwrite:
                raise ValueError("Unexpectedly tried to overwrite block data")
            self._overwrite(ids, rows)
            return
        self._seen.add(ids)
        self._id_blocks.append(ids)
        self._row_blocks.append(rows)

    def _overwrite(self, ids, rows):
        """Dict semantics for repeated ids: a known id keeps its position and takes the new row; new ids
        are appended in order (the rare path: the reference only uses it with allow_overwrite=True)."""
        merged = self.embed_data
        for i, r in zip(ids.tolist(), rows):
            merged[i] = r
        all_ids, all_rows = dict_to_arrays(merged)
        self._id_blocks, self._row_blocks = [all_ids], [all_rows]
        self._seen = _IdSet()
        self._seen.add(all_ids)

    def save_shard(self):
        """Write this rank's rows to <tmp>/<rank>.pkl — or <tmp>/<rank>.rows.f16 + .ids.i64 (:63-70)."""
        os.makedirs(self.temp_dir_name, exist_ok=True)
        base = '{}/{}'.format(self.temp_dir_name, self.rank)
        if self.format == "flat":
            ids, rows = self.to_arrays()
            save_flat(base + '.pkl', ids, rows)
            return
        with open(base + '.pkl', 'wb') as writer:
            pickle.dump(self.state(), writer)

    def merge_shards_and_save(self):
        """Fold every other rank's shard into this one, write the merged store, remove tmp (:72-100)."""
        shard_ranks = sorted({int(f.split('.')[0]) for f in os.listdir(self.temp_dir_name)})
        seen_own_shard = False
        for shard_rank in shard_ranks:
            if shard_rank == self.rank:
                seen_own_shard = True
                continue
            base = '{}/{}.pkl'.format(self.temp_dir_name, shard_rank)
            if flat_exists(base):
                ids, rows = load_flat(base)
                ids, rows = np.array(ids), np.array(rows)
            else:
                with open(base, 'rb') as f:
                    ids, rows = dict_to_arrays(pickle.load(f)['embed_data'])
            old_size = len(self)
            try:
                self._append(ids, rows, allow_overwrite=False)
            except ValueError:
                raise AssertionError("evidence shards overlap")
            assert len(self) == old_size + ids.shape[0], "evidence shards overlap"
        assert seen_own_shard
        if self.format == "flat":
            self.save_flat()
        else:
            with open(self.embedding_path, 'wb') as final_file:
                pickle.dump(self.state(), final_file)
        shutil.rmtree(self.temp_dir_name, ignore_errors=True)

    # ------------------------------------------------------------------ dense views / flat format
    def to_arrays(self):
        """(ids int64 [N], rows float16 [N, d]) in insertion order (what :245-249 builds)."""
        if not self._id_blocks:
            return np.zeros(0, dtype=np.int64), np.zeros((0, 0), dtype=np.float16)
        if len(self._id_blocks) > 1:
            self._id_blocks = [np.concatenate(self._id_blocks)]
            self._row_blocks = [np.concatenate(self._row_blocks, axis=0)]
        return self._id_blocks[0], self._row_blocks[0]

    def save_flat(self, path=None):
        ids, rows = self.to_arrays()
        return save_flat(path or self.embedding_path, ids, rows)


class _IdSet(object):
    """Membership of doc ids without a 21 M-entry Python set: a growable bitmap for the non-negative
    ids the reference uses (1-based TSV row numbers, orqa_wiki_dataset.py:192), a set for anything else."""

    LIMIT = 1 << 31

    def __init__(self):
        self.bits = np.zeros(0, dtype=bool)
        self.other = set()

    def _split(self, ids):
        small = (ids >= 0) & (ids < self.LIMIT)
        return ids[small], ids[~small]

    def any_known(self, ids):
        small, big = self._split(ids)
        inside = small[small < self.bits.shape[0]]
        if inside.shape[0] and self.bits[inside].any():
            return True
        return any(int(i) in self.other for i in big)

    def add(self, ids):
        small, big = self._split(ids)
        if small.shape[0]:
            top = int(small.max()) + 1
            if top > self.bits.shape[0]:
                grown = np.zeros(max(top, 2 * self.bits.shape[0]), dtype=bool)
                grown[:self.bits.shape[0]] = self.bits
                self.bits = grown
            self.bits[small] = True
        self.other.update(int(i) for i in big)


def dict_to_arrays(embed_data):
    """{doc_id: float16[d]} -> (ids int64 [N], rows float16 [N, d]), insertion order preserved.  One
    C-level concatenate over the values instead of a Python loop over N rows."""
    n = len(embed_data)
    ids = np.fromiter(embed_data.keys(), dtype=np.int64, count=n)
    if n == 0:
        return ids, np.zeros((0, 0), dtype=np.float16)
    values = list(embed_data.values())
    d = int(np.asarray(values[0]).shape[0])
    try:
        rows = np.concatenate(values).reshape(n, d)
    except ValueError:                       # scalars / ragged input: fall back to the generic conversion
        rows = np.asarray(values).reshape(n, d)
    if rows.dtype != np.float16:
        rows = rows.astype(np.float16)
    return ids, rows


def flat_paths(path):
    base = os.path.splitext(path)[0]
    return base + '.rows.f16', base + '.ids.i64'


def flat_exists(path):
    rows_path, ids_path = flat_paths(path)
    return os.path.exists(rows_path) and os.path.exists(ids_path)


def flat_shape(path):
    hdr = np.fromfile(flat_paths(path)[0], dtype=np.int64, count=2)
    return int(hdr[0]), int(hdr[1])


def save_flat(path, ids, rows):
    rows_path, ids_path = flat_paths(path)
    rows = np.ascontiguousarray(rows, dtype=np.float16)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    if rows.shape[0] != ids.shape[0]:
        raise ValueError("rows and ids disagree on N")
    d = rows.shape[1] if rows.ndim == 2 else 0
    with open(rows_path + '.tmp', 'wb') as f:
        f.write(np.array([rows.shape[0], d], dtype=np.int64).tobytes())
        f.write(rows.tobytes())
    with open(ids_path + '.tmp', 'wb') as f:
        f.write(ids.tobytes())
    os.replace(rows_path + '.tmp', rows_path)        # readers never see a half-written store
    os.replace(ids_path + '.tmp', ids_path)
    return rows_path, ids_path


def load_flat(path, row_range=None):
    """Memory-map the flat store; row_range=(lo, hi) returns only that slice (no full read)."""
    rows_path, ids_path = flat_paths(path)
    n, d = flat_shape(path)
    if n == 0:
        return np.zeros(0, dtype=np.int64), np.zeros((0, d), dtype=np.float16)
    rows = np.memmap(rows_path, dtype=np.float16, mode='r', offset=16, shape=(n, d))
    ids = np.memmap(ids_path, dtype=np.int64, mode='r', shape=(n,))
    if row_range is not None:
        lo, hi = row_range
        rows, ids = rows[lo:hi], ids[lo:hi]
    return ids, rows
