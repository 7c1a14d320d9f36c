"""Evidence re-encoding (index refresh) with the context tower on the B200 kernels.

Mirrors reference megatron/indexer_emdr2.py:38-114 (`IndexBuilder.build_and_save_index`): stream the
evidence set through the context tower in batches, store one float16 row per doc id, save this
rank's shard, let the main builder merge.  Two outputs are offered:

  * `build_and_save_index()`  the reference's protocol: EvidenceStore pickles on disk
    (emdr2_b200/store.py), so a reference trainer can pick the refreshed index up unchanged.
  * `build_into_index(index)` B200-native refresh: embeddings are written straight into a new
    HBM-resident shard and bound to a B200BruteForceIndex — no `.cpu()` sync per batch
    (indexer_emdr2.py:95, emdr2_index.py:12-13), no per-row dict insert, no 2 x 32 GB pickle.

The batch source is any iterable of (row_id int64 [b], context_tokens int64 [b, s],
context_types int64 [b, s]) — `get_open_retrieval_batch` (megatron/data/orqa_wiki_dataset.py) with
the dense mask dropped, since the kernels derive it from the token ids.
"""
import numpy as np
import torch

from .store import EvidenceStore


class IndexBuilder(object):
    def __init__(self, model, batches, embedding_path=None, rank=0, world=1, group=None,
                 log_interval=0):
        self.model = model                      # a DualEncoder (or anything with .context_model)
        self.batches = batches
        self.embedding_path = embedding_path
        self.rank, self.world, self.group = rank, world, group
        self.is_main_builder = rank == 0
        self.num_total_builders = world
        self.log_interval = log_interval
        self.iteration = self.total_processed = 0
        self.evidence_embedder_obj = None
        if embedding_path is not None:
            self.evidence_embedder_obj = EvidenceStore(embedding_path, load_from_path=False, rank=rank)

    def _context_tower(self):
        m = self.model
        while not hasattr(m, "context_model") and hasattr(m, "module"):
            m = m.module
        return m.context_model if hasattr(m, "context_model") else m

    @staticmethod
    def _tower_takes_lengths(tower):
        import inspect
        try:
            return "row_lengths" in inspect.signature(tower.forward).parameters
        except (TypeError, ValueError):
            return False

    def track_and_report_progress(self, batch_size):
        self.iteration += 1
        self.total_processed += batch_size * self.num_total_builders
        if self.is_main_builder and self.log_interval and self.iteration % self.log_interval == 0:
            print('Batch {:10d} | Total {:10d}'.format(self.iteration, self.total_processed), flush=True)

    def _barrier(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and self.world > 1:
            dist.barrier(group=self.group)

    def embed_batches(self):
        """Yields (row_id tensor, embeddings [b, h] 16-bit CUDA) per batch; no host sync."""
        tower = self._context_tower()
        for row_id, tokens, types in self.batches:
            dev = next(tower.parameters()).device
            kw = {}
            if not tokens.is_cuda and self._tower_takes_lengths(tower):
                # lengths are free on the host: the tower then skips all-padding columns and runs the
                # batch length-bucketed (blocks.py: encode); CLS embeddings are unchanged
                t = tokens.numpy()
                lens = ((t != 0) * np.arange(1, t.shape[1] + 1)).max(axis=1)
                kw = dict(max_len=int(lens.max()) if lens.size else 0, row_lengths=lens)
            with torch.no_grad():
                emb = tower(tokens.to(dev, non_blocking=True), None, types.to(dev, non_blocking=True), **kw)
            self.track_and_report_progress(batch_size=len(row_id))
            yield row_id, emb

    def build_and_save_index(self, expected_total=None):
        if self.evidence_embedder_obj is None:
            raise RuntimeError("build_and_save_index needs an embedding_path")
        for row_id, emb in self.embed_batches():
            self.evidence_embedder_obj.add_block_data(row_id.cpu().numpy(), emb.float().cpu().numpy())
        self.evidence_embedder_obj.save_shard()
        self._barrier()
        if self.is_main_builder:
            self.evidence_embedder_obj.merge_shards_and_save()
            if expected_total is not None:      # "every single piece of data was embedded" (:110)
                assert len(self.evidence_embedder_obj) == expected_total
        self.evidence_embedder_obj.clear()
        self._barrier()

    def build_into_index(self, index, dtype=torch.float16):
        """Encode this rank's batches and bind them as the index's resident shard (ids on device)."""
        ids, rows = [], []
        for row_id, emb in self.embed_batches():
            ids.append(row_id.to(emb.device, non_blocking=True).to(torch.int64))
            rows.append(emb.to(dtype))
        all_ids = torch.cat(ids) if ids else torch.empty(0, dtype=torch.int64)
        all_rows = torch.cat(rows) if rows else torch.empty(0, index.embed_size, dtype=dtype)
        index.add_local_shard(all_ids, all_rows, num_rows=int(all_ids.numel()) * self.world)
        return index
