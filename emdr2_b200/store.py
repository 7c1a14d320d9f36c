"""Evidence-embedding store: the reference's pickle format plus a flat, shard-addressable format.

Mirrors ``OpenRetreivalDataStore`` (reference megatron/data/emdr2_index.py:16-100): a pickled
``{'embed_data': {int doc_id: np.float16[d]}}`` whose row order is dict insertion order, built per
rank with ``add_block_data``/``save_shard`` and merged by one rank with ``merge_shards_and_save``.
The names, argument meaning and error behaviour (ValueError on overwrite, the no-overlap assert of
:88-90) are kept so the class can stand where the reference's does; the differences are that it takes
its path and rank explicitly (no global ``get_args()``) and that it can also read/write a flat layout
(``<path>.rows.f16`` = [N, d] float16 row-major, ``<path>.ids.i64`` = [N] int64) that a rank can
memory-map and slice to its own row range instead of unpickling all 32 GB.
"""
import os
import pickle
import shutil

import numpy as np


class EvidenceStore(object):
    """Serializable holder of evidence embeddings keyed by doc id (reference: OpenRetreivalDataStore)."""

    def __init__(self, embedding_path=None, load_from_path=True, rank=None):
        if embedding_path is None:
            raise ValueError("embedding_path is required (the reference reads args.embedding_path here)")
        self.embed_data = dict()
        self.embedding_path = embedding_path
        self.rank = 0 if rank is None else rank
        if load_from_path:
            self.load_from_file()
        block_data_name = os.path.splitext(self.embedding_path)[0]
        self.temp_dir_name = block_data_name + '_tmp'

    def state(self):
        return {'embed_data': self.embed_data}

    def clear(self):
        """Drop the embeddings (the index owns a copy once add_embed_data has run; :38-43)."""
        self.embed_data = dict()

    def load_from_file(self):
        """Populate from the pickle at embedding_path (:45-54)."""
        with open(self.embedding_path, 'rb') as f:
            state_dict = pickle.load(f)
        self.embed_data = state_dict['embed_data']

    def add_block_data(self, row_id, block_embeds, allow_overwrite=False):
        """Insert rows as np.float16 keyed by doc id (:56-61)."""
        for idx, embed in zip(row_id, block_embeds):
            idx = int(idx)
            if not allow_overwrite and idx in self.embed_data:
                raise ValueError("Unexpectedly tried to overwrite block data")
            self.embed_data[idx] = np.float16(embed)

    def save_shard(self):
        """Write this rank's rows to <tmp>/<rank>.pkl (:63-70)."""
        os.makedirs(self.temp_dir_name, exist_ok=True)
        with open('{}/{}.pkl'.format(self.temp_dir_name, self.rank), 'wb') as writer:
            pickle.dump(self.state(), writer)

    def merge_shards_and_save(self):
        """Fold every other rank's shard into this one, write the merged pickle, remove tmp (:72-100)."""
        shard_names = os.listdir(self.temp_dir_name)
        seen_own_shard = False
        for fname in shard_names:
            shard_rank = int(os.path.splitext(fname)[0])
            if shard_rank == self.rank:
                seen_own_shard = True
                continue
            with open('{}/{}'.format(self.temp_dir_name, fname), 'rb') as f:
                data = pickle.load(f)
            old_size = len(self.embed_data)
            shard_size = len(data['embed_data'])
            self.embed_data.update(data['embed_data'])
            assert len(self.embed_data) == old_size + shard_size, "evidence shards overlap"
        assert seen_own_shard
        with open(self.embedding_path, 'wb') as final_file:
            pickle.dump(self.state(), final_file)
        shutil.rmtree(self.temp_dir_name, ignore_errors=True)

    # ------------------------------------------------------------------ dense views / flat format
    def to_arrays(self):
        """(ids int64 [N], rows float16 [N, d]) in dict insertion order (what :245-249 builds)."""
        return dict_to_arrays(self.embed_data)

    def save_flat(self, path=None):
        ids, rows = self.to_arrays()
        return save_flat(path or self.embedding_path, ids, rows)


def dict_to_arrays(embed_data):
    """{doc_id: float16[d]} -> (ids int64 [N], rows float16 [N, d]), insertion order preserved."""
    n = len(embed_data)
    ids = np.fromiter(embed_data.keys(), dtype=np.int64, count=n)
    if n == 0:
        return ids, np.zeros((0, 0), dtype=np.float16)
    first = next(iter(embed_data.values()))
    d = int(np.asarray(first).shape[0])
    rows = np.empty((n, d), dtype=np.float16)
    for i, v in enumerate(embed_data.values()):
        rows[i] = v
    return ids, rows


def flat_paths(path):
    base = os.path.splitext(path)[0]
    return base + '.rows.f16', base + '.ids.i64'


def save_flat(path, ids, rows):
    rows_path, ids_path = flat_paths(path)
    rows = np.ascontiguousarray(rows, dtype=np.float16)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    if rows.shape[0] != ids.shape[0]:
        raise ValueError("rows and ids disagree on N")
    with open(rows_path, 'wb') as f:
        f.write(np.array([rows.shape[0], rows.shape[1]], dtype=np.int64).tobytes())
        f.write(rows.tobytes())
    with open(ids_path, 'wb') as f:
        f.write(ids.tobytes())
    return rows_path, ids_path


def load_flat(path, row_range=None):
    """Memory-map the flat store; row_range=(lo, hi) returns only that slice (no full read)."""
    rows_path, ids_path = flat_paths(path)
    hdr = np.fromfile(rows_path, dtype=np.int64, count=2)
    n, d = int(hdr[0]), int(hdr[1])
    rows = np.memmap(rows_path, dtype=np.float16, mode='r', offset=16, shape=(n, d))
    ids = np.memmap(ids_path, dtype=np.int64, mode='r', shape=(n,))
    if row_range is not None:
        lo, hi = row_range
        rows, ids = rows[lo:hi], ids[lo:hi]
    return ids, rows
