"""Dropout bookkeeping for the training path (reference: torch dropout with p = hidden_dropout /
attention_dropout = 0.1, megatron/arguments.py:218-221, applied at transformer.py:345-346,397-419,511-515
and language_model.py:181).

The kernels regenerate every mask from (seed, offset, row, column) (csrc/dropout.cuh), so all the host keeps
is a seed, a call counter that hands each dropout call site a fresh `offset`, and — per device — the seed's
column-hash table.  A `DropoutSpec` is what a forward op records for its backward: the same four numbers
give the same mask.  Ranks of a data-parallel group should seed differently (`manual_seed(seed + rank)`)."""
import ctypes
import threading

import torch

from . import _lib

TABLE_COLUMNS = 65536          # keys of the longest FiD cross-attention (50 x 512 = 25 600) fit with room


class DropoutSpec(object):
    __slots__ = ("p", "seed", "offset", "colhash")

    def __init__(self, p, seed, offset, colhash):
        self.p, self.seed, self.offset, self.colhash = float(p), int(seed), int(offset), colhash

    def c_args(self):
        return (ctypes.c_float(self.p), ctypes.c_uint64(self.seed), ctypes.c_uint64(self.offset),
                ctypes.c_void_p(self.colhash.data_ptr()))


class DropoutState(object):
    def __init__(self, seed=1234):
        self._lock = threading.Lock()
        self.manual_seed(seed)

    def manual_seed(self, seed):
        with self._lock:
            self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
            self.counter = 0
            self._tables = {}

    def table(self, device):
        device = torch.device(device)
        key = (device.type, device.index)
        t = self._tables.get(key)
        if t is None:
            t = torch.empty(TABLE_COLUMNS, dtype=torch.int32, device=device)
            with torch.cuda.device(device):
                _lib.check(_lib.load().emdr2_dropout_colhash(
                    ctypes.c_uint64(self.seed), ctypes.c_void_p(t.data_ptr()), TABLE_COLUMNS,
                    ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "emdr2_dropout_colhash")
            self._tables[key] = t
        return t

    def next(self, p, device, columns):
        """A fresh spec for one dropout call over `columns` columns (None when p == 0)."""
        if not p:
            return None
        if columns + 128 > TABLE_COLUMNS:
            raise ValueError("dropout over %d columns exceeds the column-hash table (%d)" % (columns, TABLE_COLUMNS))
        with self._lock:
            self.counter += 1
            offset = self.counter
        return DropoutSpec(p, self.seed, offset, self.table(device))


#: process-wide state used by the modules of blocks.py; re-seed with `manual_seed`
STATE = DropoutState()


def manual_seed(seed):
    STATE.manual_seed(seed)


def mask(spec, rows, cols):
    """uint8 [rows, cols]: 1 where the spec keeps element (row, col).  For tests that replay a mask."""
    out = torch.empty((rows, cols), dtype=torch.uint8, device=spec.colhash.device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().emdr2_dropout_mask(
            *spec.c_args(), ctypes.c_void_p(out.data_ptr()), rows, cols,
            ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)), "emdr2_dropout_mask")
    return out


def keep_scale(p):
    """1 / (1 - p_eff) with p_eff = round(p * 2^32) / 2^32: the factor the kernels apply to kept values."""
    return 1.0 / (1.0 - min(int(p * 4294967296.0 + 0.5), 4294967295) / 4294967296.0)
