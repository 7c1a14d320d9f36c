"""Generate tests/golden/blocks_ref_*.npz by running the REFERENCE's own BERT tower and T5 reader.

Build-container only (needs /root/reference).  The reference modules are executed unmodified on
CPU in fp32: megatron.model.dualencoder_model.PretrainedBertModel (:146-181) and
megatron.model.t5_model.T5Model (:84-154) over megatron/model/{language_model,transformer}.py.
What is patched, and only that:
  * the four import shims of SURVEY.md §8c (torch._six, apex, amp_C, np.float);
  * megatron.global_vars._GLOBAL_ARGS = a namespace holding the tiny model config;
    _GLOBAL_TOKENIZER = a stub exposing .pad (PretrainedBertModel reads tokenizer.pad, :152);
  * torch.distributed = single-rank gloo; mpu.initialize_model_parallel(1);
  * torch.cuda.current_device() -> 'cpu' (transformer.py:306 allocates the score buffer there);
  * mpu.get_cuda_rng_tracker().fork() -> null context (transformer.py:345; dropout is off: eval()).

Weights are NOT stored: both this script and the tests fill every parameter from
tests/helpers.py:seeded_weights(name, shape) so the fixtures hold only inputs and outputs.
Config (head dim 64, the only one the CUDA attention kernel supports): hidden 128, 2 heads,
2 layers, ffn 256, vocab 128, max positions 64.
"""
import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import make_mips_golden  # noqa: E402  (shims)
from helpers import TINY, seeded_weights, tiny_inputs  # noqa: E402


def setup_reference():
    make_mips_golden.install_shims()
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29541", rank=0, world_size=1)
    from megatron import global_vars, mpu
    args = types.SimpleNamespace(
        hidden_size=TINY["hidden"], num_attention_heads=TINY["heads"], num_layers=TINY["layers"],
        ffn_hidden_size=TINY["ffn"], kv_channels=TINY["hidden"] // TINY["heads"],
        max_position_embeddings=TINY["max_pos"], padded_vocab_size=TINY["vocab"],
        hidden_dropout=0.1, attention_dropout=0.1, layernorm_epsilon=1e-5, init_method_std=0.02,
        apply_query_key_layer_scaling=False, attention_softmax_in_fp32=False,
        apply_residual_connection_post_layernorm=False, bias_gelu_fusion=False,
        bias_dropout_fusion=False, scaled_masked_softmax_fusion=False,
        scaled_upper_triang_masked_softmax_fusion=False, checkpoint_activations=False,
        checkpoint_num_layers=1, num_unique_layers=None, param_sharing_style="grouped",
        openai_gelu=False, onnx_safe=None, fp16=False, fp16_lm_cross_entropy=False,
        params_dtype=torch.float32, use_cpu_initialization=True, model_parallel_size=1,
        bert_load=None, rank=0)
    global_vars._GLOBAL_ARGS = args
    global_vars._GLOBAL_TOKENIZER = types.SimpleNamespace(pad=0)
    if not mpu.model_parallel_is_initialized():
        mpu.initialize_model_parallel(1)
    torch.cuda.current_device = lambda: "cpu"
    tracker = types.SimpleNamespace(fork=lambda *a, **k: contextlib.nullcontext())
    mpu.get_cuda_rng_tracker = lambda: tracker
    import megatron.mpu.random as mrandom
    mrandom.get_cuda_rng_tracker = lambda: tracker
    import megatron.model.transformer as tr
    tr.mpu.get_cuda_rng_tracker = lambda: tracker


def fill(model):
    names = []
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.copy_(seeded_weights(name, tuple(p.shape)))
            names.append(name)
    return names


def main():
    setup_reference()
    from megatron.data.mask_creation_utils import make_attention_mask_3d, make_history_mask_3d
    from megatron.model.dualencoder_model import PretrainedBertModel
    from megatron.model.t5_model import T5Model
    inp = tiny_inputs()

    bert = PretrainedBertModel(num_tokentypes=2, parallel_output=True, vocab_size=TINY["vocab"]).eval()
    bert_names = fill(bert)
    ids, types_ = torch.from_numpy(inp["bert_ids"]), torch.from_numpy(inp["bert_types"])
    mask = make_attention_mask_3d(ids, ids) < 0.5                   # emdr2_model.py:119-120
    with torch.no_grad():
        pooled = bert(ids, mask, types_)
        hidden = bert.language_model(ids, torch.arange(ids.shape[1])[None].expand_as(ids),
                                     mask.unsqueeze(1), tokentype_ids=types_)
    np.savez_compressed(os.path.join(HERE, "blocks_ref_bert.npz"), pooled=pooled.numpy(),
                        hidden=hidden.numpy(), names=np.array(bert_names))
    print("bert: pooled", tuple(pooled.shape), "hidden", tuple(hidden.shape), len(bert_names), "params")

    t5 = T5Model(num_tokentypes=2, parallel_output=True, vocab_size=TINY["vocab"]).eval()
    t5_names = fill(t5)
    enc, dec = torch.from_numpy(inp["t5_enc_ids"]), torch.from_numpy(inp["t5_dec_ids"])
    enc_mask = make_attention_mask_3d(enc, enc) < 0.5               # emdr2_model.py:148-149
    dec_mask = (make_attention_mask_3d(dec, dec) * make_history_mask_3d(dec)) < 0.5   # :169-171
    cross_mask = make_attention_mask_3d(dec, enc) < 0.5             # :166-167
    with torch.no_grad():
        logits, enc_out = t5(enc, dec, enc_mask, dec_mask, cross_mask)
        enc_only = t5(enc, dec, enc_mask, None, None, output_enc_hidden=True)   # :152-157
        # FiD path: encoder states of K passages concatenated along the key axis (:159-183)
        b, k, s = inp["fid_shape"]
        fid_states = enc_out.reshape(b, k * s, TINY["hidden"])
        fid_ids = enc.reshape(b, k * s)
        fid_dec = dec[:b]
        fid_cross = make_attention_mask_3d(fid_dec, fid_ids) < 0.5
        fid_dmask = (make_attention_mask_3d(fid_dec, fid_dec) * make_history_mask_3d(fid_dec)) < 0.5
        fid_logits, _ = t5(fid_ids[:, :s], fid_dec, None, fid_dmask, fid_cross,
                           enc_hidden_states=fid_states)
    assert torch.equal(enc_only, enc_out)
    np.savez_compressed(os.path.join(HERE, "blocks_ref_t5.npz"), logits=logits.numpy(),
                        enc_out=enc_out.numpy(), fid_logits=fid_logits.numpy(),
                        names=np.array(t5_names))
    print("t5: logits", tuple(logits.shape), "enc_out", tuple(enc_out.shape), "fid_logits",
          tuple(fid_logits.shape), len(t5_names), "params")
    for f in ("blocks_ref_bert.npz", "blocks_ref_t5.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
