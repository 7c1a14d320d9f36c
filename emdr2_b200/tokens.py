"""Flat token store: the token arrays of N documents in one buffer plus per-document offsets.

This is the layout of the reference's memory-mapped indexed datasets (megatron/data/indexed_dataset.py:
`MMapIndexedDataset` = one `.bin` token buffer + `.idx` pointers/sizes) that back `passages_map` and
`title_map` in the retrieval tail (megatron/model/emdr2_model.py:464-466, `x[doc_id - 1]`).  Keeping it
flat lets the step hand (buffer, offsets) to the native formatter (`emdr2_format_passages_flat`), which
reads the retrieved passages in place — no per-passage array objects, `.tolist()` calls or concatenation.
`x[i]` still returns the i-th document's tokens, so the store also stands wherever an indexable map does.
"""
import numpy as np

_WIDTHS = {np.dtype(np.uint16): 2, np.dtype(np.int32): 4, np.dtype(np.int64): 8}


class FlatTokenStore(object):
    def __init__(self, tokens, offsets):
        tokens = np.asarray(tokens)
        if tokens.dtype not in _WIDTHS:
            raise TypeError("token buffer must be uint16, int32 or int64, got %s" % tokens.dtype)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        if tokens.ndim != 1 or offsets.ndim != 1 or offsets.shape[0] < 1:
            raise ValueError("tokens must be 1-D and offsets [N + 1]")
        if offsets[0] < 0 or offsets[-1] > tokens.shape[0] or (np.diff(offsets) < 0).any():
            raise ValueError("offsets must be non-decreasing and inside the token buffer")
        self.tokens = tokens                  # may be a np.memmap
        self.offsets = offsets
        self.token_bytes = _WIDTHS[tokens.dtype]

    @classmethod
    def from_arrays(cls, arrays, dtype=np.int64):
        lens = np.fromiter((len(a) for a in arrays), dtype=np.int64, count=len(arrays))
        offsets = np.zeros(len(arrays) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        tokens = np.concatenate([np.asarray(a, dtype=dtype) for a in arrays]) if len(arrays) else \
            np.zeros(0, dtype=dtype)
        return cls(tokens.astype(dtype, copy=False), offsets)

    @classmethod
    def from_indexed_dataset(cls, dataset):
        """Wrap a reference `MMapIndexedDataset` without copying its token buffer: `_index._pointers` are
        byte offsets into `_bin_buffer`, `_index._sizes` the document lengths (indexed_dataset.py:340-480)."""
        index = dataset._index
        dtype = np.dtype(index.dtype)
        tokens = np.frombuffer(dataset._bin_buffer, dtype=dtype)
        starts = np.asarray(index._pointers, dtype=np.int64) // dtype.itemsize
        sizes = np.asarray(index._sizes, dtype=np.int64)
        if len(starts) and not np.array_equal(starts[1:], (starts + sizes)[:-1]):
            raise ValueError("documents are not stored back to back; copy them with from_arrays")
        offsets = np.concatenate([starts[:1] if len(starts) else np.zeros(1, np.int64), starts + sizes])
        return cls(tokens, offsets)

    def __len__(self):
        return self.offsets.shape[0] - 1

    def __getitem__(self, i):
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self.tokens[self.offsets[i]:self.offsets[i + 1]]

    def spans(self, index):
        """(element offsets int64, lengths int32) of the documents `index` (any shape); entries < 0 give
        (0, 0) — absent documents."""
        index = np.asarray(index, dtype=np.int64)
        ok = index >= 0
        safe = np.where(ok, index, 0)
        if safe.size and (safe.max() >= len(self)):
            raise IndexError("document index %d outside the store of %d" % (int(safe.max()), len(self)))
        start = self.offsets[safe]
        length = self.offsets[safe + 1] - start
        return np.where(ok, start, 0), np.where(ok, length, 0).astype(np.int32)
