"""Retrieval-recall evaluation: the second consumer of the MIPS index (top-k accuracy of a retriever).

Mirrors the evaluation tail of reference tasks/openqa/dense_retriever/evaluation/evaluate.py:123-168
(`OpenRetrievalEvaluator.evaluate`: search the question embeddings with `FaissMIPSIndex` at
`--topk-retrievals 100`, then count for every k how many questions have an answer-bearing passage among
their k best) and the answer matching of .../evaluation/qa_validation.py:29-133 (`calculate_matches`,
`check_answer`, `has_answer`, `regex_match`) with the tokenisation of .../evaluation/tokenizers.py:153-192
(`SimpleTokenizer`: runs of letters/digits/marks, or any other single non-space character).

The index is any object with the `FaissMIPSIndex` calling convention (`search_mips_index(queries, top_k,
reconstruct=False) -> (distances fp32 numpy, ids int64 numpy)`), i.e. `emdr2_b200.index.B200FaissMIPSIndex`
— k = 100 is served exactly through its row-range refinement.  Matching is host-side string work.
"""
import collections
import unicodedata

import regex

QAMatchStats = collections.namedtuple("QAMatchStats", ["top_k_hits", "questions_doc_hits"])

_TOKEN = regex.compile(r"([\p{L}\p{N}\p{M}]+)|([^\p{Z}\p{C}])",
                       flags=regex.IGNORECASE + regex.UNICODE + regex.MULTILINE)


def simple_words(text, uncased=True):
    """SimpleTokenizer.tokenize(text).words(uncased) (tokenizers.py:153-192, :52-61)."""
    words = [m.group() for m in _TOKEN.finditer(text)]
    return [w.lower() for w in words] if uncased else words


def _normalize(text):
    return unicodedata.normalize("NFD", text)


def regex_match(text, pattern):
    """Is the regex contained in the text (qa_validation.py:124-133); a broken pattern matches nothing."""
    try:
        compiled = regex.compile(pattern, flags=regex.IGNORECASE + regex.UNICODE + regex.MULTILINE)
    except BaseException:
        return False
    return compiled.search(text) is not None


def has_answer(answers, text, match_type="string"):
    """Does the passage contain one of the answers (qa_validation.py:96-121): token-sequence containment
    for 'string', regex search for 'regex'; anything else never matches."""
    text = _normalize(text)
    if match_type == "string":
        words = simple_words(text)
        for answer in answers:
            target = simple_words(_normalize(answer))
            n = len(target)
            for i in range(0, len(words) - n + 1):
                if target == words[i:i + n]:
                    return True
    elif match_type == "regex":
        for answer in answers:
            if regex_match(text, _normalize(answer)):
                return True
    return False


def check_answer(answers, doc_ids, id2text, match_type="string"):
    """Per retrieved passage: does it bear an answer (qa_validation.py:73-93).  id2text[doc_id] is
    (text, title) like the evidence dataset's map (orqa_wiki_dataset.py:190-196)."""
    hits = []
    for doc_id in doc_ids:
        text = id2text[doc_id][0]
        hits.append(False if text is None else has_answer(answers, text, match_type))
    return hits


def top_k_hits(questions_doc_hits, n_docs):
    """top_k_hits[k-1] = number of questions with a hit among their k best (qa_validation.py:63-68)."""
    totals = [0] * n_docs
    for hits in questions_doc_hits:
        best = next((i for i, x in enumerate(hits) if x), None)
        if best is not None:
            for i in range(best, n_docs):
                totals[i] += 1
    return totals


def calculate_matches(id2text, answers, closest_docs, match_type="string"):
    """QAMatchStats(top_k_hits, questions_doc_hits) for closest_docs = [(doc_ids, scores)] per question."""
    scores = [check_answer(a, ids, id2text, match_type) for a, (ids, _) in zip(answers, closest_docs)]
    n_docs = len(closest_docs[0][0]) if closest_docs else 0
    return QAMatchStats(top_k_hits(scores, n_docs), scores)


class RecallEvaluator(object):
    """evaluate(query_embeds, answers) -> {k: accuracy} plus the per-question hit lists, like the
    reference's "top-k: xx.xx" report (evaluate.py:157-162)."""

    def __init__(self, mips_index, id2text, topk_retrievals=100, report_topk_accuracies=(1, 5, 10, 20, 50, 100),
                 match_type="string"):
        self.mips_index = mips_index
        self.id2text = id2text
        self.topk = int(topk_retrievals)
        self.report = [k for k in report_topk_accuracies if k <= self.topk]
        self.match_type = match_type

    def evaluate(self, query_embeds, answers):
        distance, topkindex = self.mips_index.search_mips_index(query_embeds, top_k=self.topk, reconstruct=False)
        closest = [(ids.tolist(), d.tolist()) for d, ids in zip(distance, topkindex)]
        stats = calculate_matches(self.id2text, answers, closest, self.match_type)
        num_rows = max(1, len(closest))
        accuracy = {k: stats.top_k_hits[k - 1] / num_rows for k in self.report}
        return accuracy, stats, closest
