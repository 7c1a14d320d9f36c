"""Asynchronous evidence-index refresh: the trainer / indexer hand-shake and the indexer loop.

Mirrors reference tasks/openqa/e2eqa/async_indexer.py:87-144 (`AsyncIndexBuilder`), the trainer side
of the same protocol in tasks/openqa/e2eqa/train_e2eqa.py:436-505 and the group / flag set-up of
megatron/mpu/initialize.py:258-281 (`init_emdr2_groups`):

  * world ranks [0, max_training_rank) train, ranks [max_training_rank, world) re-encode the evidence;
  * one Gloo group over ALL ranks carries two one-element flag tensors:
      NEW_CHKPT_READY  broadcast from trainer rank 0, blocking on both sides: "the retriever weights
                       to index with are on disk" — starts the first build (:121, train :445) and
                       ends every hand-over (:143, train :485);
      NEW_INDEX_READY  broadcast from the main indexer (rank max_training_rank), posted ASYNC by both
                       sides (:137-140, train :449,498): "a new index is ready"; the trainers poll
                       its completion (`is_completed()`, train :480) once `index_reload_interval`
                       iterations have passed and only then stop to save + reload.

`RefreshProtocol` is that state machine with the group and ranks passed explicitly (the reference
reads them from module globals).  `AsyncIndexBuilder` is the indexer loop on top of
emdr2_b200/indexer.py:IndexBuilder.  Two hand-over modes:

  "store"   the reference's: every indexer pickles its shard, the main one merges them into
            `embedding_path`, trainers call `evidence_retriever.update_evidence_embedding()` which
            reloads that file (emdr2_model.py:426-432).  Byte-compatible with a reference trainer.
  "direct"  B200-native: indexer i sends its (ids, rows) blocks straight to the
            trainer rank that owns those rows under the `torch.chunk` split; the trainer receives them
            into a STANDBY shard buffer in HBM at the hand-over point (the live shard stays valid and
            searchable until the swap) and then swaps the two.  No pickle, no host copy of the 32 GB
            matrix; the indexers keep their rows in their own HBM until the trainers are ready for them.  `ShardReceiver` /
            `send_rows_to_owners` implement the exchange over any torch.distributed backend (NCCL
            between GPUs; Gloo in the CPU tests).
"""
import time

import numpy as np
import torch
import torch.distributed as dist

from .index import chunk_range
from .indexer import IndexBuilder


class RefreshProtocol(object):
    """The two-flag hand-shake.  `group` must be a Gloo group over all `world_size` ranks (the
    reference creates it with a 4 h timeout, initialize.py:261-263)."""

    def __init__(self, rank, world_size, max_training_rank, group=None):
        if not 0 < max_training_rank < world_size:
            raise ValueError("need at least one trainer and one indexer rank "
                             "(max_training_rank=%d, world_size=%d)" % (max_training_rank, world_size))
        self.rank, self.world_size = int(rank), int(world_size)
        self.max_training_rank = int(max_training_rank)
        self.main_builder_idx = self.max_training_rank          # async_indexer.py:93
        self.group = group
        self.new_index_ready = torch.zeros(1)                   # initialize.py:269
        self.new_chkpt_ready = torch.zeros(1)                   # initialize.py:276
        self._index_handle = None
        self.last_reload_iteration = None
        self.reloads = 0

    @property
    def is_trainer(self):
        return self.rank < self.max_training_rank

    # ---------------------------------------------------------------- both sides
    def _chkpt_barrier(self):
        """Blocking broadcast of NEW_CHKPT_READY from trainer rank 0 (every rank takes part)."""
        dist.broadcast(self.new_chkpt_ready, 0, group=self.group)

    def _post_index_ready(self):
        return dist.broadcast(self.new_index_ready, self.main_builder_idx, group=self.group, async_op=True)

    # ---------------------------------------------------------------- trainer side (train_e2eqa.py)
    def trainer_start(self, iteration=0):
        """:436-453 — release the indexers for their first build and start listening for its end."""
        assert self.is_trainer
        self._chkpt_barrier()
        self._index_handle = self._post_index_ready()
        self.last_reload_iteration = iteration

    def index_is_ready(self):
        return self._index_handle is not None and self._index_handle.is_completed()

    def reload_due(self, iteration, index_reload_interval):
        return iteration >= self.last_reload_iteration + index_reload_interval

    def trainer_maybe_reload(self, iteration, index_reload_interval, save_checkpoint, update_index,
                             poll_seconds=5.0, wait=True, receive_index=None):
        """:477-505, called once per training iteration.  Returns True when a hand-over happened.
        When the reload is due the reference WAITS for the indexers (sleeping 5 s between polls);
        wait=False turns that into a non-blocking check (train on with the old index).
        receive_index (direct hand-over only) runs BEFORE the checkpoint barrier: it is the receiving
        end of the indexers' point-to-point sends, which they complete before joining the barrier."""
        assert self.is_trainer
        if not self.reload_due(iteration, index_reload_interval):
            return False
        while not self.index_is_ready():
            if not wait:
                return False
            time.sleep(poll_seconds)
        save_checkpoint(iteration)                 # the weights the next index is built with
        if receive_index is not None:
            receive_index()
        self._chkpt_barrier()                      # ... are on disk: indexers may load them
        update_index()                             # evidence_retriever.update_evidence_embedding()
        self._index_handle = self._post_index_ready()
        self.last_reload_iteration = iteration
        self.reloads += 1
        return True

    # ---------------------------------------------------------------- indexer side (async_indexer.py)
    def indexer_wait_for_start(self):
        """:121 — block until trainer rank 0 announces the first checkpoint."""
        assert not self.is_trainer
        self._chkpt_barrier()

    def indexer_announce_index(self, deliver=None):
        """:131-143 — announce the finished index (async), then block until the trainers have saved
        the checkpoint the next build starts from.  deliver (direct hand-over only) sends the rows to
        their owners in between: the trainers start receiving once they have seen the announcement."""
        assert not self.is_trainer
        self._post_index_ready()
        if deliver is not None:
            deliver()
        self._chkpt_barrier()


# ------------------------------------------------------------------------------- direct hand-over
def owner_of_rows(num_rows, num_trainers):
    """Trainer rank that owns each global row under torch.chunk (emdr2_index.py:252)."""
    size = -(-num_rows // num_trainers) if num_rows else 1
    return np.minimum(np.arange(num_rows) // size, num_trainers - 1)


def send_rows_to_owners(row_index, ids, rows, num_rows, num_trainers, group=None, trainer_ranks=None):
    """Indexer side of the direct hand-over: `row_index` (int64 [n], global row numbers in evidence
    order), `ids` (int64 [n] doc ids) and `rows` ([n, d] embeddings) of this indexer are cut by owner
    and sent with one point-to-point message triple per trainer: header (count), then row numbers +
    ids, then the rows.  Every indexer sends to every trainer (possibly a count of 0), so the
    receivers know when they are done."""
    trainer_ranks = list(range(num_trainers)) if trainer_ranks is None else list(trainer_ranks)
    row_index = torch.as_tensor(row_index, dtype=torch.int64)
    ids = torch.as_tensor(ids, dtype=torch.int64)
    dev = rows.device
    size = -(-num_rows // num_trainers) if num_rows else 1
    owner = torch.clamp(row_index // size, max=num_trainers - 1)
    for t, dst in enumerate(trainer_ranks):
        sel = torch.nonzero(owner == t).reshape(-1)
        header = torch.tensor([int(sel.numel())], dtype=torch.int64, device=dev)
        dist.send(header, dst, group=group)
        if sel.numel() == 0:
            continue
        meta = torch.stack((row_index[sel], ids[sel])).to(dev).contiguous()
        dist.send(meta, dst, group=group)
        dist.send(rows[sel.to(dev)].contiguous(), dst, group=group)


class ShardReceiver(object):
    """Trainer side of the direct hand-over: a standby [n_local, d] shard (+ ids) that indexer
    messages are scattered into while the live shard keeps serving searches; `swap_into(index)`
    makes it the live one (B200BruteForceIndex.add_local_shard) and the next refresh reuses the
    retired buffer — two resident copies of the local shard, 2 x 4 GB per GPU at 21 M x 768 / 8."""

    def __init__(self, num_rows, embed_size, trainer_rank, num_trainers, dtype=torch.float16, device="cpu"):
        self.num_rows, self.embed_size = int(num_rows), int(embed_size)
        self.rank, self.world = int(trainer_rank), int(num_trainers)
        self.row_lo, self.row_hi = chunk_range(self.num_rows, self.world, self.rank)
        n_local = self.row_hi - self.row_lo
        self.dtype, self.device = dtype, torch.device(device)
        self.buffers = [None, None]
        self.standby = 0
        self._alloc(0, n_local)
        self.filled = 0

    def _alloc(self, i, n_local):
        if self.buffers[i] is None:
            self.buffers[i] = (torch.empty((n_local, self.embed_size), dtype=self.dtype, device=self.device),
                               torch.full((n_local,), -1, dtype=torch.int64, device=self.device))
        return self.buffers[i]

    def receive_from(self, indexer_ranks, group=None):
        """Blocks until every indexer has delivered its part of this trainer's row range."""
        rows_buf, ids_buf = self._alloc(self.standby, self.row_hi - self.row_lo)
        ids_buf.fill_(-1)
        self.filled = 0
        for src in indexer_ranks:
            header = torch.zeros(1, dtype=torch.int64, device=self.device)
            dist.recv(header, src, group=group)
            n = int(header.item())
            if n == 0:
                continue
            meta = torch.empty((2, n), dtype=torch.int64, device=self.device)
            dist.recv(meta, src, group=group)
            part = torch.empty((n, self.embed_size), dtype=self.dtype, device=self.device)
            dist.recv(part, src, group=group)
            local = meta[0] - self.row_lo
            if int(local.min()) < 0 or int(local.max()) >= rows_buf.shape[0]:
                raise ValueError("indexer %d sent rows outside [%d, %d)" % (src, self.row_lo, self.row_hi))
            rows_buf[local] = part
            ids_buf[local] = meta[1]
            self.filled += n
        if self.filled != rows_buf.shape[0] or bool((ids_buf < 0).any()):
            # "make sure that every single piece of data was embedded" (indexer_emdr2.py:110)
            raise RuntimeError("refresh incomplete: %d of %d local rows received" % (self.filled, rows_buf.shape[0]))
        return rows_buf, ids_buf

    def swap_into(self, index):
        """Make the standby shard the index's live shard; the retired one becomes the next standby."""
        rows_buf, ids_buf = self.buffers[self.standby]
        index.add_local_shard(ids_buf, rows_buf, num_rows=self.num_rows, row_lo=self.row_lo)
        self.standby = 1 - self.standby


# --------------------------------------------------------------- same GPUs, second stream (config 5)
class ConcurrentShardRefresher(object):
    """Index refresh CONCURRENT with training on the same GPUs (BASELINE config 5; the reference needs a second
    set of 8 GPUs for it, README.md:113-115, async_indexer.py:116-144).

    Every trainer rank re-encodes ITS OWN row range of the evidence — rows are independent, so the refresh shards
    exactly like the index and nothing has to be sent anywhere: a worker thread drives the context
    tower on a side CUDA stream (a private, frozen copy of the weights taken at `start`: the checkpoint the
    reference's indexers load, :127-129) and writes the embeddings straight into a STANDBY shard buffer in this
    GPU's HBM, while the training thread keeps stepping on the main stream and the live shard keeps serving
    searches.  `maybe_swap()` is called by the training thread between two steps: the ranks agree (one tiny
    all-reduce of ready flags — the NEW_INDEX_READY of the reference's protocol) that every standby is complete,
    and only then do ALL of them swap, so a collective search never mixes old and new shards.  The retired
    buffer becomes the next standby; writes to it are ordered after the searches that may still read it.

    tower           the context tower (BertTower-like: tower(tokens, None, types, max_len=, row_lengths=) -> [b, d])
    make_batches()  fresh iterable of (row_id int64 [b] — 1-based doc ids of rows in [row_lo, row_hi) —,
                    tokens int64 [b, s], types int64 [b, s]) covering this rank's range (pinned host tensors)
    partial=True    rows no batch covers are copied from the live shard (bounded benchmark runs)
    """

    def __init__(self, index, tower, make_batches, group=None, partial=False):
        import copy
        self.index, self.group, self.partial = index, group, partial
        self.make_batches = make_batches
        self.device = index.device
        self.live_tower = tower
        self.tower = copy.deepcopy(tower).eval()
        for p in self.tower.parameters():
            p.requires_grad_(False)
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self.standby = None
        self._retired_free = None          # event: the retired buffer's last reader has been enqueued before it
        self._thread = None
        self._done = None
        self.error = None
        self.rows_done = 0
        self.seconds = None
        self.rounds = 0

    def start(self):
        """Snapshot the context-tower weights and start re-encoding on the side stream."""
        import threading
        if self._thread is not None and self._thread.is_alive():
            raise RuntimeError("a refresh is already running")
        n_local, d = self.index.evidence_embeds.shape
        if self.standby is None:
            self.standby = torch.empty((n_local, d), dtype=self.index.dtype, device=self.device)
        self.rows_done, self.error, self.seconds = 0, None, None
        fork = None
        if self.stream is not None:
            fork = torch.cuda.Event()
            fork.record(torch.cuda.current_stream(self.device))      # weights as of this point of the training stream
        self._thread = threading.Thread(target=self._run, args=(fork,), daemon=True)
        self._thread.start()

    def _run(self, fork):
        import contextlib
        import time as _time
        try:
            t0 = _time.perf_counter()
            if self.stream is not None:
                torch.cuda.set_device(self.device)                   # a new thread starts on device 0
            ctx = torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()
            with ctx, torch.no_grad():
                if self.stream is not None:
                    self.stream.wait_event(fork)
                    if self._retired_free is not None:
                        self.stream.wait_event(self._retired_free)
                for dst, src in zip(self.tower.parameters(), self.live_tower.parameters()):
                    dst.copy_(src)                                   # the "checkpoint" the indexer works from
                if self.partial:
                    self.standby.copy_(self.index.evidence_embeds)
                lo = self.index.row_lo
                takes_lengths = IndexBuilder._tower_takes_lengths(self.tower)
                for row_id, tokens, types in self.make_batches():
                    kw = {}
                    if takes_lengths and not tokens.is_cuda:
                        t = tokens.numpy()
                        lens = ((t != 0) * np.arange(1, t.shape[1] + 1)).max(axis=1)
                        kw = dict(max_len=int(lens.max()) if lens.size else 0, row_lengths=lens)
                    emb = self.tower(tokens.to(self.device, non_blocking=True), None,
                                     types.to(self.device, non_blocking=True), **kw)
                    rows = torch.as_tensor(row_id, dtype=torch.int64).to(self.device, non_blocking=True) - 1 - lo
                    self.standby.index_copy_(0, rows, emb.to(self.standby.dtype))
                    self.rows_done += int(len(row_id))
                if self.stream is not None:
                    self._done = torch.cuda.Event()
                    self._done.record(self.stream)
                    self._done.synchronize()                         # this THREAD waits; the training thread does not
            self.seconds = _time.perf_counter() - t0
        except BaseException as exc:                                 # surfaced by ready() / maybe_swap()
            self.error = exc

    def ready(self):
        if self.error is not None:
            raise RuntimeError("index refresh failed") from self.error
        return self._thread is not None and not self._thread.is_alive()

    def maybe_swap(self):
        """Training thread, between two steps.  True when every rank's standby was complete and all swapped."""
        if self._thread is None:
            return False
        flag = torch.tensor([1.0 if self.ready() else 0.0], device=self.device)
        if self.group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if float(flag.item()) < 1.0:
            return False
        self._thread.join()
        self._thread = None
        (old_ids, old_rows), free = self.index.swap_local_shard(self.index.local_ids, self.standby)
        self.standby, self._retired_free = old_rows, free
        self.rounds += 1
        return True


# ----------------------------------------------------------------------------------- indexer loop
class AsyncIndexBuilder(IndexBuilder):
    """The indexer process (async_indexer.py:87-144): wait for the first checkpoint, then forever
    {re-encode the evidence, hand the index over, load the context-tower weights the trainers just
    saved}.

    make_batches()      -> fresh iterable of (row_id, tokens, types) for THIS indexer's share of the
                           evidence (get_one_epoch_dataloader over the indexer group, indexer_emdr2.py:16-35)
    load_weights(model) -> loads 'retriever/biencoder_model' context weights from the trainers' last
                           checkpoint (load_attributes(custom_load_path=args.load, ...), :127-129)
    mode "store": EvidenceStore pickles at `embedding_path`; mode "direct": rows go to the owning
    trainers (`num_rows` global rows, row numbers = doc id - 1, the TSV order of orqa_wiki_dataset.py:192).
    """

    def __init__(self, model, make_batches, protocol, load_weights=None, embedding_path=None, index_rank=0,
                 index_world=1, index_group=None, mode="store", num_rows=None, data_group=None, log_interval=0):
        if mode not in ("store", "direct"):
            raise ValueError("mode must be 'store' or 'direct'")
        if mode == "direct" and num_rows is None:
            raise ValueError("direct hand-over needs the global number of evidence rows")
        super().__init__(model, None, embedding_path=embedding_path if mode == "store" else None,
                         rank=index_rank, world=index_world, group=index_group, log_interval=log_interval)
        self.make_batches = make_batches
        self.protocol = protocol
        self.load_weights = load_weights
        self.mode, self.num_rows, self.data_group = mode, num_rows, data_group
        self.rounds = 0
        self._pending = None

    def build_once(self):
        self.batches = self.make_batches()
        self.iteration = self.total_processed = 0
        if self.mode == "store":
            self.build_and_save_index()
            return
        ids, rows = [], []
        for row_id, emb in self.embed_batches():
            ids.append(torch.as_tensor(row_id, dtype=torch.int64).reshape(-1))
            rows.append(emb.to(torch.float16))
        all_ids = torch.cat(ids) if ids else torch.empty(0, dtype=torch.int64)
        all_rows = torch.cat(rows) if rows else torch.empty((0, 1), dtype=torch.float16)
        self._pending = (all_ids, all_rows)        # stays in this indexer's memory until the hand-over

    def _deliver(self):
        all_ids, all_rows = self._pending
        self._pending = None
        send_rows_to_owners(all_ids - 1, all_ids, all_rows, self.num_rows, self.protocol.max_training_rank,
                            group=self.data_group)

    def run_async(self, max_rounds=None):
        """:116-129.  max_rounds bounds the otherwise endless loop (tests, benchmarks)."""
        self.protocol.indexer_wait_for_start()
        while max_rounds is None or self.rounds < max_rounds:
            if self.is_main_builder:
                print("Starting Indexing again!", flush=True)
            self.build_once()
            self.protocol.indexer_announce_index(deliver=self._deliver if self.mode == "direct" else None)
            if self.load_weights is not None:
                self.load_weights(self.model)
            self.rounds += 1
