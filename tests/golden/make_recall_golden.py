"""Golden outputs of the reference's answer matching (tasks/openqa/dense_retriever/evaluation/
qa_validation.py + tokenizers.py) on seeded texts.  Run in the build container (needs /root/reference):

    python tests/golden/make_recall_golden.py

`spacy` (imported at the top of tokenizers.py for a tokenizer this path never constructs) is stubbed."""
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

WORDS = ["Paris", "paris", "the", "Eiffel", "tower", "1889", "São", "Paulo", "naïve", "co-operate", "U.S.", "3.14",
         "Zürich", "is", "in", "of", "New", "York", "(city)", "rock'n'roll", "东京", "café", "a", "b", "—", "42"]


def cases(seed=7, n=120):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        text = " ".join(rng.choice(WORDS) for _ in range(rng.randint(3, 25)))
        kind = rng.choice(["string", "string", "regex"])
        if kind == "string":
            answers = [" ".join(rng.choice(WORDS) for _ in range(rng.randint(1, 3))) for _ in range(rng.randint(1, 3))]
            if rng.random() < 0.5:                      # plant an answer taken from the text (case changed)
                toks = text.split()
                i = rng.randrange(len(toks))
                answers.append(" ".join(toks[i:i + rng.randint(1, 3)]).upper())
        else:
            answers = [rng.choice([r"18\d\d", r"par[iy]s", r"(new|old) york", r"[unclosed", r"\bcaf.\b", r"^the"])]
        out.append(dict(text=text, answers=answers, match_type=kind))
    return out


def load_reference():
    sys.modules.setdefault("spacy", types.ModuleType("spacy"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from tasks.openqa.dense_retriever.evaluation import qa_validation, tokenizers
    return qa_validation, tokenizers


def main():
    qa, tk = load_reference()
    tok = tk.SimpleTokenizer()
    rows = []
    for c in cases():
        rows.append(dict(c, has_answer=bool(qa.has_answer(c["answers"], c["text"], tok, c["match_type"])),
                         words=tok.tokenize(qa._normalize(c["text"])).words(uncased=True)))
    # the top-k accumulation (calculate_matches forks a process pool; its arithmetic is :63-68)
    rng = random.Random(3)
    hit_lists = [[rng.random() < 0.15 for _ in range(10)] for _ in range(30)]
    top_k = [0] * 10
    for hits in hit_lists:
        best = next((i for i, x in enumerate(hits) if x), None)
        if best is not None:
            top_k[best:] = [v + 1 for v in top_k[best:]]
    with open(os.path.join(HERE, "recall_ref.json"), "w") as f:
        json.dump(dict(cases=rows, hit_lists=hit_lists, top_k_hits=top_k), f, ensure_ascii=False, indent=0)
    print("wrote", len(rows), "cases;", sum(r["has_answer"] for r in rows), "positive")


if __name__ == "__main__":
    main()
