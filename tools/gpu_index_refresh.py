"""Evidence re-encoding rate (BASELINE configs[4], one GPU's share): NQ-shaped passages (S = 256, indexer
batch 128, reference arguments.py:589) through the context tower into a resident shard
(IndexBuilder.build_into_index), timed with CUDA events.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emdr2_b200.blocks import bert_base_config
from emdr2_b200.index import B200BruteForceIndex
from emdr2_b200.indexer import IndexBuilder
from emdr2_b200.model import DualEncoder

DEV = "cuda:0"
BATCH, SEQ, BATCHES = int(os.environ.get("BATCH", "128")), 256, int(os.environ.get("BATCHES", "60"))
torch.manual_seed(0)
model = DualEncoder(bert_base_config(torch.bfloat16), bert_vocab_size=30592, only_context_model=True).to(DEV).eval()
with torch.no_grad():
    for name, p in model.named_parameters():
        p.normal_(0.0, 0.02) if p.dim() > 1 else p.zero_()
        if "layernorm" in name and name.endswith("weight"):
            p.fill_(1.0)
rng = np.random.RandomState(0)


def batches(n):
    for i in range(n):
        ids = np.zeros((BATCH, SEQ), dtype=np.int64)
        for r, ln in enumerate(rng.randint(105, 192, size=BATCH)):
            ids[r, :ln] = rng.randint(1000, 30000, size=ln)
        yield torch.arange(i * BATCH + 1, (i + 1) * BATCH + 1), torch.from_numpy(ids).pin_memory(), \
            torch.zeros((BATCH, SEQ), dtype=torch.int64).pin_memory()


index = B200BruteForceIndex(768, device=DEV)
IndexBuilder(model, list(batches(3))).build_into_index(index)          # warm-up
torch.cuda.synchronize()
data = list(batches(BATCHES))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
IndexBuilder(model, data).build_into_index(index)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
rate = BATCH * BATCHES / (ms * 1e-3)
print(json.dumps({"metric": "passages/s re-encoded into a resident shard (context tower, BERT-base, S=256)",
                  "value": rate, "unit": "passages/s", "batch": BATCH, "batches": BATCHES, "ms": ms,
                  "full_refresh_21M_s_on_8_gpus": 21e6 / rate / 8,
                  "rows_bound": int(index.row_hi - index.row_lo)}))
