"""Golden cases for the passage formatter and the losses, produced by the REFERENCE's own functions
(megatron/model/emdr2_model.py:306-376 query_*_t5_format, megatron/data/orqa_wiki_dataset.py:86-120,
tasks/openqa/e2eqa/train_e2eqa.py:72-123,184-214).  Build-container only; outputs committed."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_mips_golden  # noqa: E402


def random_case(rng):
    n_docs = int(rng.choice([1, 2, 3]))
    docs = [rng.randint(5, 100, size=int(rng.randint(1, 30))).tolist() for _ in range(n_docs)]
    main = int(rng.choice({1: [0], 2: [0, -1], 3: [1]}[n_docs]))
    if n_docs == 2 and rng.rand() < 0.3:
        main = 1 if rng.rand() < 0.5 else main      # reference also reaches the "middle" branch with 2 docs
    return dict(query=rng.randint(5, 100, size=int(rng.randint(1, 12))).tolist(),
                title=rng.randint(5, 100, size=int(rng.randint(1, 6))).tolist(),
                docs=docs, main=main, max_len=int(rng.choice([24, 32, 48, 64])))


def main():
    make_mips_golden.install_shims()
    from megatron.model.emdr2_model import query_extended_context_t5_format, query_single_context_t5_format
    from megatron.data.orqa_wiki_dataset import build_tokens_types_paddings_from_ids
    from tasks.openqa.e2eqa.train_e2eqa import get_loss_and_retriever_utility, get_kl_div_retriever
    rng = np.random.RandomState(77)
    cases = []
    for _ in range(200):
        c = random_case(rng)
        c["extended"] = query_extended_context_t5_format(list(c["query"]), list(c["title"]), [list(d) for d in c["docs"]],
                                                         c["main"], c["max_len"], 3, 0)
        c["single"] = query_single_context_t5_format(list(c["query"]), list(c["title"]), list(c["docs"][c["main"]]),
                                                     c["max_len"], 3, 0)
        ids, types, mask = build_tokens_types_paddings_from_ids(c["title"] + [3] + c["docs"][c["main"]], c["max_len"], 2, 3, 0)
        c["bert"] = [list(map(int, ids)), list(map(int, types)), mask.tolist()]
        cases.append(c)
    with open(os.path.join(HERE, "formatter_ref.json"), "w") as f:
        json.dump(cases, f)
    g = torch.Generator().manual_seed(5)
    b, k, l, v = 3, 4, 6, 50
    logits = torch.randn(b, k, l, v, generator=g) * 2
    topk_log_probs = torch.log_softmax(torch.randn(b, k, generator=g), dim=1)
    labels = torch.randint(1, v, (b, l), generator=g)
    loss_mask = (torch.rand(b, l, generator=g) < 0.7).float()
    loss_mask[:, 0] = 1
    labels = labels.masked_fill(loss_mask == 0, -1).clamp(min=-1)
    labels[0, 0] = 5
    eos_id = 40
    lm_loss, ru, null_loss = get_loss_and_retriever_utility(logits, topk_log_probs, labels.clone(), loss_mask, eos_id)
    kl = get_kl_div_retriever(logits, topk_log_probs, labels.clone(), loss_mask)
    np.savez_compressed(os.path.join(HERE, "losses_ref.npz"), logits=logits.numpy(), topk_log_probs=topk_log_probs.numpy(),
                        labels=labels.numpy(), loss_mask=loss_mask.numpy(), eos_id=np.int64(eos_id),
                        lm_loss=lm_loss.numpy(), retriever_utility=ru.numpy(), null_block_lm_loss=null_loss.numpy(),
                        kl=kl.numpy())
    print("formatter cases:", len(cases), "losses:", float(lm_loss), float(ru), float(null_loss), float(kl))


if __name__ == "__main__":
    main()
