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
