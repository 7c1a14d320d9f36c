"""Retriever façade: the collective part of ``PreComputedEvidenceDocsRetriever`` on the B200 index.

Mirrors reference megatron/model/emdr2_model.py:379-470.  The reference gathers every trainer's
queries (C1, :439), lets the node-first rank search an index that lives on all GPUs of the node
(:441-446), then broadcasts fp16 scores and int32 ids back (C4, :451-452) and slices each rank's
rows (:454-455).  Here every rank of the MIPS group owns one row range of the evidence matrix, so
after the same query all-gather each rank scans its own shard, and one all-gather of [nq, k]
(score, id) pairs + an on-device merge replaces the score-slab copies and both broadcasts
(emdr2_b200/index.py).  The return value keeps the reference's shape:
``(topk_data, distance)`` with ``topk_data[b] = (ids_K, [(doc_list, main_doc_idx, title_ids)]*K)``.
"""
import torch

from .index import B200BruteForceIndex
from .store import EvidenceStore


class B200EvidenceRetriever(object):
    """topk / embedding_size / embedding_path / allow_trivial_doc follow the reference's args
    (arguments.py:583,593; emdr2_model.py:382-391): without ``allow_trivial_doc`` one extra
    passage is fetched so the caller can drop the gold one.

    passages_map / title_map are indexable token stores (``x[doc_id - 1] -> np.ndarray``) and
    wikititledocmap has ``get_neighbour_paragraphs(doc_id) -> (doc_ids, main_idx)``
    (tools/inverted_title_index.py:22-37); when they are omitted ``get_topk`` returns ids only.
    """

    index_cls = B200BruteForceIndex

    def __init__(self, topk, embedding_size, embedding_path=None, allow_trivial_doc=True,
                 group=None, dtype=torch.float16, passages_map=None, title_map=None,
                 wikititledocmap=None, store=None):
        self.topk = int(topk) + (0 if allow_trivial_doc else 1)
        self.allow_trivial_doc = allow_trivial_doc
        self.embedding_size = int(embedding_size)
        self.group = group
        self.passages_map = passages_map
        self.title_map = title_map
        self.wikititledocmap = wikititledocmap
        self.evidence_embedder_obj = store
        if store is None and embedding_path is not None:
            self.get_evidence_embedding(embedding_path)
        self.mips_index = self.index_cls(embed_size=self.embedding_size,
                                         embed_data=self.evidence_embedder_obj, group=group,
                                         dtype=dtype)
        self._barrier()

    def get_evidence_embedding(self, path):
        self.evidence_embedder_obj = EvidenceStore(path, load_from_path=True)

    def _barrier(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=self.group)

    def update_evidence_embedding(self):
        """Reload the refreshed index from disk on every rank (:426-432)."""
        self.mips_index.update_index()
        self._barrier()

    def search_all(self, query_tensor):
        """All-gather the local queries [B, d] over the group and search: returns
        (scores fp32 [W*B, k], ids int64 [W*B, k]) — identical on every rank."""
        import torch.distributed as dist
        q = query_tensor.detach().contiguous()
        world = self.mips_index.world
        if world > 1:
            allq = torch.empty((world * q.shape[0], q.shape[1]), dtype=q.dtype, device=q.device)
            dist.all_gather_into_tensor(allq, q, group=self.group)
        else:
            allq = q
        return self.mips_index.search(allq, self.topk)

    #: EMDR2Model passes as_arrays=True when the retriever advertises it (skips ~B*K*4 .tolist() calls)
    supports_arrays = True

    @property
    def supports_packed(self):
        """True when the token maps are flat stores and the title map answers whole batches
        (emdr2_b200/tokens.py:FlatTokenStore, titlemap.py:NeighbourTable or anything with `lookup`):
        `get_topk(..., as_packed=True)` then does the tail of emdr2_model.py:457-468 with a handful of
        numpy gathers and hands the formatter offsets instead of token arrays."""
        from .tokens import FlatTokenStore
        return isinstance(self.passages_map, FlatTokenStore) and isinstance(self.title_map, FlatTokenStore) \
            and hasattr(self.wikititledocmap, "lookup")

    def get_topk_packed(self, query_tensor):
        """(PackedTopk, distance): the retrieval tail without touching a token."""
        import numpy as np
        from .formatter import PackedTopk
        local_bsize = query_tensor.shape[0]
        scores, ids = self.search_all(query_tensor)
        rank = self.mips_index.rank
        mine = slice(rank * local_bsize, (rank + 1) * local_bsize)
        distance = scores[mine].to(torch.float16)
        topk = ids[mine].to(torch.int32).cpu().numpy().astype(np.int64)       # ONE device->host copy
        bsz, k = topk.shape
        flat_ids = topk.reshape(-1)
        docs, n_docs, main_idx = self.wikititledocmap.lookup(flat_ids)          # [n, 3], [n], [n]
        t_off, t_len = self.title_map.spans(flat_ids - 1)
        d_off, d_len = self.passages_map.spans(np.where(docs >= 0, docs - 1, -1))
        meta = np.empty((flat_ids.shape[0], 6), dtype=np.int32)
        meta[:, 0], meta[:, 1], meta[:, 2] = t_len, n_docs, main_idx
        meta[:, 3:] = d_len
        piece = np.empty((flat_ids.shape[0], 4), dtype=np.int64)
        piece[:, 0] = t_off
        piece[:, 1:] = d_off
        cand_begin = np.arange(bsz + 1, dtype=np.int32) * k
        return PackedTopk(cand_begin, flat_ids, meta, piece, self.title_map, self.passages_map), distance

    def get_topk(self, query_tensor, as_arrays=False, as_packed=False):
        """(topk_data, distance) like the reference.  as_arrays=True keeps the passage / title tokens
        as the int64 arrays the token store returns instead of converting them to Python lists
        (emdr2_model.py:464-466 calls .tolist() on each); formatter.postprocess takes either."""
        if as_packed:
            return self.get_topk_packed(query_tensor)
        local_bsize = query_tensor.shape[0]
        scores, ids = self.search_all(query_tensor)
        rank = self.mips_index.rank
        mine = slice(rank * local_bsize, (rank + 1) * local_bsize)
        distance = scores[mine].to(torch.float16)
        topkindex = ids[mine].to(torch.int32)
        rows = topkindex.tolist()            # ONE device->host copy (the reference does B*K .item()s)
        topk_data = []
        for topkarray in rows:
            if self.passages_map is None:
                topk_data.append((topkarray, None))
                continue
            text_list = []
            for idx in topkarray:
                doc_idxs, main_doc_idx = self.wikititledocmap.get_neighbour_paragraphs(idx)
                if as_arrays:
                    doc_list = [self.passages_map[doc_id - 1] for doc_id in doc_idxs]
                    text_list.append((doc_list, main_doc_idx, self.title_map[idx - 1]))
                else:
                    doc_list = [self.passages_map[doc_id - 1].tolist() for doc_id in doc_idxs]
                    text_list.append((doc_list, main_doc_idx, self.title_map[idx - 1].tolist()))
            topk_data.append((topkarray, text_list))
        return topk_data, distance
