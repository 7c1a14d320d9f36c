"""Retrieval-recall evaluation (emdr2_b200/recall.py): answer matching against outputs of the reference's
qa_validation.has_answer / SimpleTokenizer (tests/golden/recall_ref.json, made by make_recall_golden.py),
and the evaluator end to end at k = 100 over the oracle-backed index double (CPU)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from emdr2_b200 import recall

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
sys.path.insert(0, HERE)


def _golden():
    with open(os.path.join(GOLDEN, "recall_ref.json")) as f:
        return json.load(f)


def test_matching_and_tokenisation_equal_the_reference_fixture():
    g = _golden()
    assert len(g["cases"]) >= 100
    for c in g["cases"]:
        assert recall.simple_words(recall._normalize(c["text"])) == c["words"], c["text"]
        assert recall.has_answer(c["answers"], c["text"], c["match_type"]) == c["has_answer"], c
    assert recall.top_k_hits(g["hit_lists"], 10) == g["top_k_hits"]
    assert recall.has_answer(["x"], "x y", "unknown-match-type") is False


@pytest.mark.skipif(not os.path.isdir("/root/reference/tasks"), reason="reference not mounted")
def test_matching_equals_the_live_reference_on_fresh_cases():
    sys.path.insert(0, GOLDEN)
    import make_recall_golden as mk
    qa, tk = mk.load_reference()
    tok = tk.SimpleTokenizer()
    for c in mk.cases(seed=99, n=300):
        assert recall.has_answer(c["answers"], c["text"], c["match_type"]) == \
            bool(qa.has_answer(c["answers"], c["text"], tok, c["match_type"])), c


def test_recall_evaluator_top100_over_the_index_double():
    """Questions whose embedding is close to one passage's: that passage holds the answer, so accuracy
    must be 1.0 from the rank at which the oracle places it; k = 100 goes through the exact
    row-range refinement of the index."""
    from test_host_logic import CpuDoubleIndex
    from emdr2_b200.index import B200FaissMIPSIndex
    from oracle import mips as oracle

    class FaissDouble(B200FaissMIPSIndex):
        searcher_factory = staticmethod(CpuDoubleIndex.searcher_factory)
        merge_fn = staticmethod(CpuDoubleIndex.merge_fn)

    rng = np.random.RandomState(2)
    n, d, nq, k = 1500, 16, 6, 100
    rows = (rng.randint(-127, 128, size=(n, d)) / 64).astype(np.float16)
    ids = np.arange(1, n + 1, dtype=np.int64)
    gold = rng.choice(n, size=nq, replace=False)
    queries = (rows[gold].astype(np.float32) * 2 + rng.randint(-8, 9, size=(nq, d)) / 64).astype(np.float16)
    id2text = {int(i): ("filler text number %d" % i, "title") for i in ids}
    answers = []
    for qi, row in enumerate(gold):
        id2text[int(ids[row])] = ("the answer is Zürich-%d indeed" % qi, "title")
        answers.append(["zürich-%d" % qi])
    index = FaissDouble(d, device="cpu")
    index.add_arrays(ids, rows)
    ev = recall.RecallEvaluator(index, id2text, topk_retrievals=k, report_topk_accuracies=(1, 5, 20, 100, 200))
    acc, stats, closest = ev.evaluate(torch.from_numpy(queries), answers)
    want_s, want_i = oracle.mips_topk(rows, queries, k, ids=ids)
    assert sorted(acc) == [1, 5, 20, 100] and acc[100] == 1.0
    for qi in range(nq):
        rank = int(np.where(want_i[qi] == ids[gold[qi]])[0][0])
        assert stats.questions_doc_hits[qi].index(True) == rank
        assert closest[qi][0] == want_i[qi].tolist()
    for kk in (1, 5, 20, 100):
        want = sum(int(np.where(want_i[qi] == ids[gold[qi]])[0][0]) < kk for qi in range(nq)) / nq
        assert acc[kk] == want
