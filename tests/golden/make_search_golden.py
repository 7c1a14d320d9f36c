"""Golden outputs of the reference's decoding strategies (megatron/model/search_strategy.py) driven by a
deterministic stand-in reader on CPU.  Run in the build container (needs /root/reference):

    python tests/golden/make_search_golden.py

The module is loaded by file path (it imports only collections/numpy/torch); its hard-coded `.cuda()`
calls are made no-ops for the duration of the run.  Nothing else of the reference is touched."""
import importlib.util
import json
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/megatron/model/search_strategy.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_search_strategy", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class StubReader(object):
    """EMDR2Model stand-in: the next-token logits depend on the question, the last token and the
    position, so beams of different questions diverge and the cached per-row state must follow its
    hypothesis through every re-ordering for the output to come out right."""

    def __init__(self, vocab, eos_id, seed):
        g = torch.Generator().manual_seed(seed)
        self.vocab, self.eos_id = vocab, eos_id
        self.table = torch.randn(vocab, vocab, generator=g) * 2.0
        self.pos = torch.randn(64, vocab, generator=g)
        self.qproj = torch.randn(8, vocab, generator=g)
        self.calls = 0

    def __call__(self, query_uid, query_ids_bert, query_types, query_mask_bert, query_ids_t5, query_ids_t5_len,
                 dec_ids, all_query_context_hidden_states=None, all_query_context_ids_unflat=None,
                 topk_log_probs=None):
        self.calls += 1
        if all_query_context_hidden_states is None:
            feat = query_ids_bert[:, :8].float() / 7.0
            all_query_context_hidden_states = (feat @ self.qproj)[:, None, :]              # [B, 1, V]
            all_query_context_ids_unflat = query_ids_bert[:, :3].clone()
            topk_log_probs = torch.log_softmax(feat[:, :2], dim=1)
        rows, t = dec_ids.shape
        assert all_query_context_hidden_states.shape[0] == rows
        assert all_query_context_ids_unflat.shape[0] == rows and topk_log_probs.shape[0] == rows
        last = self.table[dec_ids[:, -1]] + self.pos[t] + all_query_context_hidden_states[:, 0, :]
        last[:, self.eos_id] += 1.5 * t - 3.5          # EOS becomes likely after a few steps
        logits = torch.zeros(rows, t, self.vocab)
        logits[:, -1, :] = last
        return logits, topk_log_probs, all_query_context_hidden_states, all_query_context_ids_unflat


def question_batch(batch, seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randint(1, 50, (batch, 12), generator=g)
    return (-torch.arange(1, batch + 1), q, torch.zeros_like(q), None, q[:, :6].clone(),
            torch.full((batch,), 6, dtype=torch.int64))


CONFIGS = [dict(batch=b, vocab=v, beam=k, alpha=a, max_len=m, seed=s)
           for (b, v, k, a, m, s) in [(1, 11, 1, 0.6, 10, 1), (4, 17, 1, 0.6, 12, 2), (3, 13, 2, 0.6, 9, 3),
                                      (5, 23, 4, 0.6, 14, 4), (2, 9, 5, 1.0, 16, 5), (6, 31, 3, 0.0, 6, 6),
                                      (4, 19, 5, 0.6, 3, 7), (7, 29, 2, 0.3, 20, 8)]]


def run(mod, cfg):
    bos, eos = 1, cfg["vocab"] - 1
    model = StubReader(cfg["vocab"], eos, cfg["seed"])
    inputs = question_batch(cfg["batch"], 100 + cfg["seed"])
    if cfg["beam"] == 1:
        obj = mod.SampleOrGreedySearch(max_decode_len=cfg["max_len"], bos_id=bos, eos_id=eos, sample=False,
                                       topk_evidence=2)
    else:
        obj = mod.BeamSearch(max_decode_len=cfg["max_len"], bos_id=bos, eos_id=eos, beam_size=cfg["beam"],
                             alpha=cfg["alpha"], topk_evidence=2)
    out = obj.generate_output(model, *inputs)
    return [[int(t) for t in row] for row in out], model.calls


def main():
    mod = load_reference()
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda()
    try:
        cases = []
        for cfg in CONFIGS:
            out, calls = run(mod, cfg)
            cases.append(dict(cfg=cfg, out=out, calls=calls))
    finally:
        torch.Tensor.cuda = cuda
    with open(os.path.join(HERE, "search_ref.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("wrote %d cases" % len(cases), [len(c["out"]) for c in cases])


if __name__ == "__main__":
    main()
