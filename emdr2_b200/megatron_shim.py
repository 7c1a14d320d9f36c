"""The reference's module surface over the B200 implementation: same class names, same constructor
signatures, configuration read from a `get_args()` namespace — so the reference's own entry points
(`tasks/openqa/e2eqa/run.py:32-38` model_provider, `train_e2eqa.py:126-181` forward step, `:556` train)
run on this package's modules unchanged.

    reference symbol (file:line)                                        here
    megatron.model.EMDR2Model(evidence_retriever)  emdr2_model.py:31    EMDR2Model
    megatron.model.PreComputedEvidenceDocsRetriever()        :379       PreComputedEvidenceDocsRetriever
    megatron.model.T5Model(num_tokentypes, parallel_output, vocab_size) t5_model.py:84      T5Model
    PretrainedBertModel(num_tokentypes, parallel_output, vocab_size)    dualencoder_model.py:146  PretrainedBertModel
    DualEncoderModel / dualencoder_model_provider            dualencoder_model.py:27,14    same names
    DistributedBruteForceIndex / FaissMIPSIndex / OpenRetreivalDataStore  data/emdr2_index.py:200,103,16

Arguments come from `get_args()` of this module: `set_args(namespace)` installs one explicitly; otherwise
the reference's own `megatron.get_args()` is used when `megatron` is importable (the drop-in case: this
package imported inside the reference's tree).  `install_into_megatron()` rebinds the names inside the
reference's modules — the two-line change a maintainer makes instead of editing call sites.
Tokenizer-derived constants (vocabulary sizes, [CLS]/[SEP]/pad ids) come from `megatron.get_tokenizer()` /
`get_t5_tokenizer()` when present, else from the namespace (`bert_vocab_size`, `t5_vocab_size`, `cls_id`,
`sep_id`, `pad_id`).  Mixed precision: `args.fp16` -> float16 parameters (the reference's only 16-bit
mode); `args.bf16` or `args.params_dtype` select bfloat16.
"""
import torch

from . import blocks, model as _model
from .index import B200BruteForceIndex, B200FaissMIPSIndex
from .retriever import B200EvidenceRetriever
from .store import EvidenceStore

_ARGS = None
_TOKENIZERS = {}


def set_args(namespace):
    """Install the argument namespace (what megatron.global_vars.set_global_variables parses)."""
    global _ARGS
    _ARGS = namespace


def set_tokenizers(bert=None, t5=None):
    _TOKENIZERS["bert"], _TOKENIZERS["t5"] = bert, t5


def get_args():
    if _ARGS is not None:
        return _ARGS
    try:
        from megatron import get_args as megatron_get_args
        return megatron_get_args()
    except Exception:
        raise RuntimeError("no arguments: call emdr2_b200.megatron_shim.set_args(namespace) "
                           "(or initialise megatron's global variables)")


def _tokenizer(kind):
    tok = _TOKENIZERS.get(kind)
    if tok is not None:
        return tok
    try:
        import megatron
        return megatron.get_tokenizer() if kind == "bert" else megatron.get_t5_tokenizer()
    except Exception:
        return None


def vocab_size_with_padding(num_tokens, args):
    """Pad the vocabulary so that it divides make_vocab_size_divisible_by * model_parallel_size
    (megatron/model/utils.py:72-79)."""
    multiple = getattr(args, "make_vocab_size_divisible_by", 128) * getattr(args, "model_parallel_size", 1)
    return -(-int(num_tokens) // multiple) * multiple


def _vocab(kind, args):
    explicit = getattr(args, kind + "_vocab_size", None)
    if explicit is not None:
        return int(explicit)
    tok = _tokenizer(kind)
    if tok is not None:
        return vocab_size_with_padding(tok.vocab_size, args)
    return int(args.padded_vocab_size)


def params_dtype(args):
    dt = getattr(args, "params_dtype", None)
    if dt in (torch.float16, torch.bfloat16):
        return dt
    if getattr(args, "bf16", False):
        return torch.bfloat16
    return torch.float16           # the reference's 16-bit mode (--fp16); there is no fp32 kernel path


def config_from_args(args=None):
    """The per-model constants the reference reads in get_language_model / ParallelTransformer
    (language_model.py:45-66, transformer.py:566-600)."""
    args = args or get_args()
    if getattr(args, "model_parallel_size", 1) != 1:
        raise ValueError("tensor model parallelism is asserted off on this path (dualencoder_model.py:15)")
    hidden = int(args.hidden_size)
    return dict(hidden=hidden, heads=int(args.num_attention_heads), layers=int(args.num_layers),
                ffn=int(getattr(args, "ffn_hidden_size", None) or 4 * hidden), vocab=int(getattr(args, "padded_vocab_size", 0) or 0),
                max_pos=int(args.max_position_embeddings), eps=float(getattr(args, "layernorm_epsilon", 1e-5)),
                dtype=params_dtype(args), hidden_dropout=float(getattr(args, "hidden_dropout", 0.1)),
                attention_dropout=float(getattr(args, "attention_dropout", 0.1)),
                dropout_seed=int(getattr(args, "seed", 1234)))


def settings_from_args(args=None):
    """What EMDR2Model.forward and postprocess read from get_args() / the tokenizers on every call
    (emdr2_model.py:92,103-104,130-131,250-303)."""
    args = args or get_args()
    bert = _tokenizer("bert")
    return dict(topk_retrievals=int(args.topk_retrievals), seq_length=int(args.seq_length),
                seq_length_ret=int(args.seq_length_ret),
                retriever_score_scaling=bool(getattr(args, "retriever_score_scaling", False)),
                update_retriever=bool(getattr(args, "update_retriever", False)),
                no_query_embedder_training=bool(getattr(args, "no_query_embedder_training", False)),
                no_context_embedder_training=bool(getattr(args, "no_context_embedder_training", False)),
                disable_retriever_dropout=bool(getattr(args, "disable_retriever_dropout", False)),
                cls_id=int(bert.cls if bert is not None else getattr(args, "cls_id", 101)),
                sep_id=int(bert.sep if bert is not None else getattr(args, "sep_id", 102)),
                pad_id=int(bert.pad if bert is not None else getattr(args, "pad_id", 0)))


# ------------------------------------------------------------------------------------------- models
class T5Model(blocks.T5Reader):
    def __init__(self, num_tokentypes=2, parallel_output=True, vocab_size=None):
        args = get_args()
        super().__init__(config_from_args(args), num_tokentypes=num_tokentypes,
                         vocab_size=vocab_size or _vocab("t5", args))
        self.parallel_output = parallel_output
        self._language_model_key, self._lm_head_key = "language_model", "lm_head"


class PretrainedBertModel(blocks.BertTower):
    def __init__(self, num_tokentypes=2, parallel_output=True, vocab_size=None):
        args = get_args()
        super().__init__(config_from_args(args), num_tokentypes=num_tokentypes,
                         vocab_size=vocab_size or _vocab("bert", args))
        self.parallel_output = parallel_output
        self._language_model_key = "language_model"


class DualEncoderModel(_model.DualEncoder):
    def __init__(self, num_tokentypes=2, parallel_output=True, only_query_model=False,
                 only_context_model=False, vocab_size=None):
        args = get_args()
        super().__init__(config_from_args(args), bert_vocab_size=vocab_size or _vocab("bert", args),
                         only_query_model=only_query_model, only_context_model=only_context_model)
        self._query_key, self._context_key = "query_model", "context_model"


def dualencoder_model_provider(only_query_model=False, only_context_model=False, vocab_size=None):
    """dualencoder_model.py:14-24."""
    return DualEncoderModel(num_tokentypes=2, parallel_output=True, only_query_model=only_query_model,
                            only_context_model=only_context_model, vocab_size=vocab_size)


class EMDR2Model(_model.EMDR2Model):
    """EMDR2Model(evidence_retriever): configuration from get_args() like emdr2_model.py:31-61."""

    def __init__(self, evidence_retriever):
        args = get_args()
        super().__init__(config_from_args(args), evidence_retriever, settings_from_args(args),
                         t5_vocab_size=_vocab("t5", args), bert_vocab_size=_vocab("bert", args))
        self.bert_tokenizer, self.t5_tokenizer = _tokenizer("bert"), _tokenizer("t5")


# ---------------------------------------------------------------------------------- retrieval side
class OpenRetreivalDataStore(EvidenceStore):
    """OpenRetreivalDataStore(embedding_path=None, load_from_path=True, rank=None): path and rank default
    to args.embedding_path / args.rank (emdr2_index.py:20-31)."""

    def __init__(self, embedding_path=None, load_from_path=True, rank=None, format="pickle"):
        if embedding_path is None or rank is None:
            args = get_args()
            embedding_path = embedding_path or args.embedding_path
            rank = getattr(args, "rank", 0) if rank is None else rank
        super().__init__(embedding_path, load_from_path=load_from_path, rank=rank, format=format)


DistributedBruteForceIndex = B200BruteForceIndex
FaissMIPSIndex = B200FaissMIPSIndex


def _data_parallel_group():
    try:
        from megatron import mpu
        return mpu.get_data_parallel_group()
    except Exception:
        import torch.distributed as dist
        return dist.group.WORLD if dist.is_available() and dist.is_initialized() else None


class PreComputedEvidenceDocsRetriever(B200EvidenceRetriever):
    """No-argument constructor like emdr2_model.py:379-406: top-k, embedding size and path, the trivial-doc
    switch and the three evidence maps all come from get_args().  The maps are built with the reference's
    own readers when they are importable (make_indexed_dataset, memory-mapped) and wrapped as flat token
    stores without copying; a namespace may also carry ready objects (`passages_map`, `title_map`,
    `wikititledocmap`).  Every rank of the data-parallel group owns one row range of the index."""

    def __init__(self):
        args = get_args()
        passages, titles, doc_map = (getattr(args, n, None) for n in ("passages_map", "title_map", "wikititledocmap"))
        if passages is None and getattr(args, "indexed_evidence_data_path", None):
            from megatron.data.indexed_dataset import make_indexed_dataset
            from .tokens import FlatTokenStore
            passages = FlatTokenStore.from_indexed_dataset(make_indexed_dataset(
                args.indexed_evidence_data_path, impl=args.data_impl, skip_warmup=(not args.mmap_warmup)))
            titles = FlatTokenStore.from_indexed_dataset(make_indexed_dataset(
                args.indexed_title_data_path, impl=args.data_impl, skip_warmup=(not args.mmap_warmup)))
        if doc_map is None and getattr(args, "evidence_data_path", None):
            from .titlemap import NeighbourTable, TitleDocMap
            doc_map = NeighbourTable(TitleDocMap(args.evidence_data_path))
        store = getattr(args, "evidence_store", None)
        super().__init__(args.topk_retrievals, args.hidden_size,
                         embedding_path=None if store is not None else getattr(args, "embedding_path", None),
                         allow_trivial_doc=bool(getattr(args, "allow_trivial_doc", False)),
                         group=_data_parallel_group(), dtype=torch.float16, passages_map=passages,
                         title_map=titles, wikititledocmap=doc_map, store=store)
        self.local_rank = getattr(args, "local_rank", 0)

    def precomputed_index_wrapper(self):
        """Kept for callers that re-run it (:408-417): (re)build the index from the store on disk."""
        self.mips_index.reset_index()
        self._barrier()


def model_provider():
    """tasks/openqa/e2eqa/run.py:32-38."""
    return EMDR2Model(PreComputedEvidenceDocsRetriever())


def install_into_megatron():
    """Rebind the reference's names to the classes above (call once after megatron is imported and its
    global variables are set).  Returns the list of (module, name) pairs that were replaced."""
    import importlib
    table = {
        "megatron.model": ["EMDR2Model", "PreComputedEvidenceDocsRetriever", "T5Model"],
        "megatron.model.emdr2_model": ["EMDR2Model", "PreComputedEvidenceDocsRetriever", "T5Model",
                                       "dualencoder_model_provider", "DistributedBruteForceIndex",
                                       "OpenRetreivalDataStore"],
        "megatron.model.t5_model": ["T5Model"],
        "megatron.model.dualencoder_model": ["PretrainedBertModel", "DualEncoderModel", "dualencoder_model_provider"],
        "megatron.data.emdr2_index": ["DistributedBruteForceIndex", "FaissMIPSIndex", "OpenRetreivalDataStore"],
        "tasks.openqa.e2eqa.run": ["EMDR2Model", "PreComputedEvidenceDocsRetriever"],
    }
    done = []
    for mod_name, names in table.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            continue
        for name in names:
            if hasattr(mod, name):
                setattr(mod, name, globals()[name])
                done.append((mod_name, name))
    return done
