"""Title -> neighbouring-passage map used by the retrieval tail.

Mirrors reference tools/inverted_title_index.py:16-66 (`WikiTitleDocMap`): passages of one article
carry the same title and consecutive doc ids; `get_neighbour_paragraphs(doc_id)` returns up to three
doc ids around `doc_id` and where the passage sits among them (0 first, 1 middle, -1 last), exactly
as emdr2_model.py:457-468 expects.  Built from the evidence TSV (id, text, title) or from
(doc_id, title) pairs.
"""
import bisect
import csv
from collections import defaultdict

import numpy as np


class TitleDocMap(object):
    def __init__(self, datapath=None, pairs=None):
        self.title2docs = defaultdict(list)
        self.docid2title = {}
        if datapath is not None:
            with open(datapath) as tsvfile:
                reader = csv.reader(tsvfile, delimiter='\t')
                next(reader, None)                       # header row
                pairs = ((int(row[0]), row[2]) for row in reader)
                self._ingest(pairs)
        elif pairs is not None:
            self._ingest(pairs)

    def _ingest(self, pairs):
        for doc_id, title in pairs:
            if doc_id in self.docid2title:
                raise AssertionError("doc id %d listed twice" % doc_id)
            self.title2docs[title].append(doc_id)
            self.docid2title[doc_id] = title

    def get_neighbour_paragraphs(self, doc_id):
        docs = self.title2docs[self.docid2title[doc_id]]
        i = bisect.bisect_left(docs, doc_id)
        if i == len(docs) or docs[i] != doc_id:
            raise ValueError(doc_id)
        if i == 0:
            return docs[0:3], 0
        if i == len(docs) - 1:
            return docs[i - 2:i + 1], -1
        return docs[i - 1:i + 2], 1


class NeighbourTable(object):
    """`get_neighbour_paragraphs` for a whole batch of doc ids at once: three dense arrays built once from
    a `TitleDocMap` (position of every passage inside its article, article length, where the article's id
    list starts in one flat array), so the 400 lookups of a step are a handful of numpy gathers instead
    of 400 dict + bisect calls.  Same results, including the reference's slice quirk for two-passage
    articles (`docs[i-2:i+1]` with i = 1 is `docs[-1:2]`: the passage alone, tools/inverted_title_index.py:34)."""

    def __init__(self, titlemap):
        ids = np.fromiter(titlemap.docid2title.keys(), dtype=np.int64, count=len(titlemap.docid2title))
        self.max_id = int(ids.max()) if ids.size else 0
        self.pos = np.full(self.max_id + 1, -1, dtype=np.int32)       # index of the passage in its article
        self.length = np.zeros(self.max_id + 1, dtype=np.int32)       # passages in its article
        self.base = np.zeros(self.max_id + 1, dtype=np.int64)         # start of the article in `flat`
        flat = []
        for docs in titlemap.title2docs.values():
            d = np.asarray(docs, dtype=np.int64)
            self.pos[d] = np.arange(d.size, dtype=np.int32)
            self.length[d] = d.size
            self.base[d] = len(flat)
            flat.extend(docs)
        self.flat = np.asarray(flat, dtype=np.int64)

    def lookup(self, doc_ids):
        """doc_ids [n] -> (neighbours int64 [n, 3] padded with -1, n_docs int32 [n], main_idx int32 [n])."""
        ids = np.asarray(doc_ids, dtype=np.int64).reshape(-1)
        if ids.size and (ids.min() < 0 or ids.max() > self.max_id or (self.pos[ids] < 0).any()):
            raise ValueError("unknown doc id in %s" % ids[:8].tolist())
        i, n, base = self.pos[ids].astype(np.int64), self.length[ids].astype(np.int64), self.base[ids]
        first = i == 0
        last = (~first) & (i == n - 1)
        alone = last & (i < 2)                                         # the slice quirk: the passage alone
        start = np.where(first, 0, np.where(last, np.maximum(i - 2, 0), i - 1))
        start = np.where(alone, i, start)
        count = np.where(first, np.minimum(n, 3), np.where(alone, 1, 3))
        cols = np.arange(3)[None, :]
        take = base[:, None] + start[:, None] + cols
        valid = cols < count[:, None]
        docs = np.where(valid, self.flat[np.where(valid, take, 0)], -1)
        main = np.where(first, 0, np.where(last, -1, 1)).astype(np.int32)
        return docs, count.astype(np.int32), main
