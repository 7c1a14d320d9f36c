"""Drop-in surface (emdr2_b200/megatron_shim.py): the reference's class names and constructor signatures
over this package's modules, configuration from a get_args() namespace.

CPU tests.  The first two need nothing but this repo; the third imports the REFERENCE's own
`_cross_entropy_forward_step` (tasks/openqa/e2eqa/train_e2eqa.py:126-181) and the reference's `model_provider`
names, and runs them unmodified on a model built by the shim (towers replaced by the CPU stand-ins of
test_model_orchestration.py — the kernels' numerics are covered by the -m gpu tests; what is under test here is
that the reference's loop can drive these modules: constructor calls, the forward signature, the three
return values, the losses computed from them).  It is skipped where /root/reference is not mounted."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from emdr2_b200 import losses, megatron_shim as shim

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLDEN = os.path.join(HERE, "golden")

H, V, K, S_RET, S, L = 64, 128, 3, 24, 64, 5


def _args(**over):
    ns = types.SimpleNamespace(
        hidden_size=H, num_attention_heads=1, num_layers=2, ffn_hidden_size=None, max_position_embeddings=64,
        padded_vocab_size=V, layernorm_epsilon=1e-5, hidden_dropout=0.1, attention_dropout=0.1, seed=1234,
        fp16=True, model_parallel_size=1, make_vocab_size_divisible_by=128, topk_retrievals=K, seq_length=S,
        seq_length_ret=S_RET, retriever_score_scaling=True, update_retriever=True, ret_kldiv=False,
        no_query_embedder_training=False, no_context_embedder_training=True, disable_retriever_dropout=False,
        allow_trivial_doc=True, embedding_path=None, rank=0, local_rank=0, async_indexer=False,
        bert_vocab_size=V, t5_vocab_size=V, cls_id=2, sep_id=3, pad_id=0)
    for k, v in over.items():
        setattr(ns, k, v)
    return ns


def test_constructors_take_their_configuration_from_get_args():
    shim.set_args(_args())
    try:
        t5 = shim.T5Model(num_tokentypes=2, parallel_output=True, vocab_size=256)
        assert t5.language_model.embedding.word_embeddings.weight.shape == (256, H)
        assert t5.language_model.add_decoder and len(t5.language_model.encoder.layers) == 2
        assert t5.lm_head.bias.dtype == torch.float16
        bert = shim.PretrainedBertModel(num_tokentypes=2, parallel_output=True)
        assert bert.language_model.embedding.word_embeddings.weight.shape == (V, H)
        dual = shim.dualencoder_model_provider(only_context_model=True)
        assert dual.use_context_model and not dual.use_query_model and not hasattr(dual, "query_model")
        model = shim.EMDR2Model(evidence_retriever=object())
        assert model.topk == K and model.settings["no_context_embedder_training"] is True
        assert model.settings["cls_id"] == 2 and model.settings["seq_length_ret"] == S_RET
        assert model._language_model_key == "encoder/t5_model" and model._retriever_model_key == "retriever/biencoder_model"
        with pytest.raises(ValueError):
            shim.config_from_args(_args(model_parallel_size=2))
        assert shim.vocab_size_with_padding(30522, _args()) == 30592       # BERT's vocabulary, padded to 128
        shim.set_args(_args(bf16=True))
        assert shim.T5Model().lm_head.bias.dtype == torch.bfloat16
    finally:
        shim.set_args(None)
    with pytest.raises(RuntimeError):
        if "megatron" not in sys.modules:
            shim.get_args()
        else:
            raise RuntimeError("megatron imported by an earlier test")


def test_saved_checkpoint_has_the_reference_nesting_and_round_trips():
    """state_dict_for_save_checkpoint writes what the reference's load_state_dict indexes
    (t5_model.py:156-176, language_model.py:367-430, dualencoder_model.py:84-109, emdr2_model.py:217-231):
    nested by 'language_model' / 'embedding' / 'encoder' / 'decoder' / 'lm_head' and 'query_model' /
    'context_model', with the reference's own parameter names at the leaves."""
    shim.set_args(_args())
    try:
        a = shim.EMDR2Model(evidence_retriever=None)
        b = shim.EMDR2Model(evidence_retriever=None)
    finally:
        shim.set_args(None)
    with torch.no_grad():
        for i, p in enumerate(a.parameters()):
            p.copy_(torch.full_like(p, 0.001 * (i + 1)))
    sd = a.state_dict_for_save_checkpoint()
    assert sorted(sd) == ["encoder/t5_model", "retriever/biencoder_model"]
    t5 = sd["encoder/t5_model"]
    assert sorted(t5) == ["language_model", "lm_head"] and list(t5["lm_head"]) == ["bias"]
    lm = t5["language_model"]
    assert sorted(lm) == ["decoder", "embedding", "encoder"]
    assert sorted(lm["embedding"]) == ["position_embeddings", "tokentype_embeddings", "word_embeddings"]
    assert list(lm["embedding"]["word_embeddings"]) == ["weight"]
    assert "layers.0.self_attention.query_key_value.weight" in lm["encoder"]
    assert "layers.1.inter_attention.key_value.bias" in lm["decoder"] and "final_layernorm.weight" in lm["decoder"]
    ret = sd["retriever/biencoder_model"]
    assert sorted(ret) == ["context_model", "query_model"] and list(ret["query_model"]) == ["language_model"]
    # leaves carry exactly the reference's named_parameters (fixture written by the reference's own modules)
    with np.load(os.path.join(GOLDEN, "blocks_ref_t5.npz")) as z:
        ref_names = sorted(str(n) for n in z["names"])

    def leaves(node, prefix=""):
        out = []
        for k, v in node.items():
            out += leaves(v, prefix + k + ".") if isinstance(v, dict) else [prefix + k]
        return out

    assert sorted(leaves(t5)) == ref_names
    b.load_state_dict(sd)                                   # nested layout in
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(p, q), n
    b.retriever_model.load_state_dict(ret)                  # what load_dualencoder_checkpoint hands over
    b.language_model.load_state_dict(t5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/megatron"), reason="reference not mounted")
def test_reference_forward_step_and_model_provider_run_on_the_shim_model():
    sys.path.insert(0, GOLDEN)
    import make_blocks_golden
    saved_current_device = torch.cuda.current_device
    try:
        make_blocks_golden.setup_reference()                # 4 import shims, 1-rank gloo, mpu, tiny args
        _drive_the_reference_step()
    finally:
        torch.cuda.current_device = saved_current_device
        shim.set_args(None)
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def _drive_the_reference_step():
    from megatron import global_vars
    from megatron.global_vars import Timers
    import tasks.openqa.e2eqa.train_e2eqa as ref_train
    import tasks.openqa.e2eqa.run as ref_run           # noqa: F401  (the module model_provider lives in)
    from test_model_orchestration import RecordingDual, RecordingReader, _setup

    ns = global_vars._GLOBAL_ARGS
    for k, v in vars(_args(padded_vocab_size=50, bert_vocab_size=50, t5_vocab_size=50)).items():
        if not hasattr(ns, k) or k in ("hidden_size", "num_attention_heads", "num_layers", "ffn_hidden_size",
                                       "max_position_embeddings", "padded_vocab_size"):
            setattr(ns, k, v)
    ns.max_training_rank = 1
    global_vars._GLOBAL_TIMERS = Timers()
    global_vars._GLOBAL_T5_TOKENIZER = types.SimpleNamespace(eos_token_id=49, vocab_size=50)
    global_vars._GLOBAL_TOKENIZER = types.SimpleNamespace(pad=0, cls=2, sep=3, vocab_size=50)

    replaced = shim.install_into_megatron()
    assert ("megatron.model", "EMDR2Model") in replaced and ("tasks.openqa.e2eqa.run", "EMDR2Model") in replaced
    import megatron.model
    assert megatron.model.EMDR2Model is shim.EMDR2Model

    # the reference's provider body (run.py:36-37) with its retriever constructor fed by the namespace
    _, _, inputs, _ = _setup(True)
    donor, log, inputs, _ = _setup(True)
    ns.passages_map, ns.title_map = donor.evidence_retriever.passages_map, donor.evidence_retriever.title_map
    ns.wikititledocmap = donor.evidence_retriever.wikititledocmap
    model = megatron.model.EMDR2Model(donor.evidence_retriever)         # reference signature: one argument
    assert model.settings["sep_id"] == 3 and model.topk == K
    model.retriever_model, model.language_model = RecordingDual(log), RecordingReader(log)
    model.train()

    uid, q_bert, q_types, _, q_t5, q_len, dec = inputs
    labels = dec.roll(-1, dims=1)
    labels[:, -1] = 0
    mask = (labels > 0).float()
    batch = dict(query_uid=uid, query_ids_bert=q_bert, query_types=q_types, query_mask_bert=torch.ones_like(q_bert),
                 query_ids_t5=q_t5, query_ids_t5_len=q_len, dec_ids=dec, labels=labels, loss_mask=mask,
                 reference=[["x"]] * uid.shape[0])
    orig_cuda, orig_sync = torch.Tensor.cuda, torch.cuda.synchronize
    torch.Tensor.cuda = lambda self, *a, **k: self          # process_batch / FloatTensor([0]).cuda() on a CPU box
    torch.cuda.synchronize = lambda *a, **k: None           # the reference's Timers bracket with a device sync
    try:
        net_loss, reduced = ref_train._cross_entropy_forward_step(batch, model)
    finally:
        torch.Tensor.cuda, torch.cuda.synchronize = orig_cuda, orig_sync
    # the same three outputs through this package's loss functions
    lm_logits, topk_log_probs, one_ctx = model(uid, q_bert, q_types, None, q_t5, q_len, dec)
    want_lm = losses.reader_cross_entropy(lm_logits, labels, mask) if lm_logits.is_cuda else \
        (torch.nn.functional.cross_entropy(lm_logits.float().view(-1, lm_logits.shape[-1]), labels.view(-1),
                                           reduction="none", ignore_index=0) * mask.view(-1)).sum() / mask.sum()
    want_ret = ref_train.get_loss_and_retriever_utility(one_ctx, topk_log_probs, labels, mask, 49)[0]
    assert torch.allclose(net_loss, want_lm + want_ret, atol=1e-6)
    assert abs(float(reduced["lm_loss"]) - float(want_lm)) < 1e-6
    assert abs(float(reduced["retriever_loss"]) - float(want_ret)) < 1e-6
    assert one_ctx.shape[:2] == (uid.shape[0], K)
