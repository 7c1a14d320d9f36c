"""Pin the block oracle (oracle/blocks.py) against the reference's own BERT tower and T5 reader
run on CPU in fp32 (tests/golden/make_blocks_golden.py -> blocks_ref_{bert,t5}.npz)."""
import os

import numpy as np
import torch

from helpers import GOLDEN, TINY, seeded_weights, tiny_inputs
from oracle import blocks


def load(name):
    with np.load(os.path.join(GOLDEN, "blocks_ref_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


def weights_for(golden, shapes):
    return {str(n): seeded_weights(str(n), shapes(str(n))) for n in golden["names"]}


def bert_shape(name):
    h, f, v, p = TINY["hidden"], TINY["ffn"], TINY["vocab"], TINY["max_pos"]
    table = {"word_embeddings.weight": (v, h), "position_embeddings.weight": (p, h),
             "tokentype_embeddings.weight": (2, h), "query_key_value.weight": (3 * h, h),
             "query_key_value.bias": (3 * h,), "key_value.weight": (2 * h, h), "key_value.bias": (2 * h,),
             "query.weight": (h, h), "query.bias": (h,), "dense.weight": (h, h), "dense.bias": (h,),
             "dense_h_to_4h.weight": (f, h), "dense_h_to_4h.bias": (f,), "dense_4h_to_h.weight": (h, f),
             "dense_4h_to_h.bias": (h,), "lm_head.bias": (v,)}
    for suffix, shape in table.items():
        if name.endswith(suffix):
            return shape
    assert "layernorm" in name, name
    return (h,)


def test_bert_tower_matches_reference():
    g = load("bert")
    w = weights_for(g, bert_shape)
    inp = tiny_inputs()
    ids, types = torch.from_numpy(inp["bert_ids"]), torch.from_numpy(inp["bert_types"])
    hidden = blocks.bert_hidden(ids, types, w, TINY["heads"], TINY["layers"])
    assert torch.allclose(hidden, torch.from_numpy(g["hidden"]), rtol=1e-5, atol=1e-5)
    pooled = blocks.bert_pooled(ids, types, w, TINY["heads"], TINY["layers"])
    assert torch.allclose(pooled, torch.from_numpy(g["pooled"]), rtol=1e-5, atol=1e-5)


def test_t5_reader_matches_reference():
    g = load("t5")
    w = weights_for(g, bert_shape)
    inp = tiny_inputs()
    enc, dec = torch.from_numpy(inp["t5_enc_ids"]), torch.from_numpy(inp["t5_dec_ids"])
    logits, enc_out = blocks.t5_forward(enc, dec, w, TINY["heads"], TINY["layers"])
    assert torch.allclose(enc_out, torch.from_numpy(g["enc_out"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(logits, torch.from_numpy(g["logits"]), rtol=1e-5, atol=2e-5)
    b, k, s = inp["fid_shape"]
    fid = blocks.t5_decode(dec[:b], enc_out.reshape(b, k * s, -1), enc.reshape(b, k * s), w,
                           TINY["heads"], TINY["layers"])
    assert torch.allclose(fid, torch.from_numpy(g["fid_logits"]), rtol=1e-5, atol=2e-5)


def test_masks_match_the_reference_definitions():
    ids = torch.tensor([[5, 3, 0, 0], [1, 2, 3, 4]])
    m = blocks.pad_mask_3d(ids, ids)
    assert m[0, 0].tolist() == [False, False, True, True] and m[0, 2].all() and not m[1].any()
    d = blocks.decoder_self_mask(ids)
    assert d[1].tolist() == torch.ones(4, 4).triu(1).bool().tolist()
