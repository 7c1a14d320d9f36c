"""In-batch-negative training step of the dual encoder (supervised retriever training): the second
consumer of the BERT towers beside the EMDR2 step.

Mirrors reference tasks/openqa/dense_retriever/train_dense_retriever.py:89-196
(`_cross_entropy_forward_step`): every rank embeds its questions and contexts (positives first, optional
hard negatives after them), the embeddings of all ranks are all-gathered — the local slice keeps its
gradient (:131-153) — scores are `Q · Cᵀ` (optionally / sqrt(hidden), :160-161), the label of question i
of rank r is the position of its positive context in the gathered list (:163-175), the loss is the mean
NLL of the row-wise log-softmax times the data-parallel world size (:178-190), and the number of
questions whose best-scoring context is their positive is reported beside it.

The towers run on the library kernels (emdr2_b200/model.py:DualEncoder); what lives here is the exchange
and the loss, written over whatever device the embeddings are on.
"""
import math

import torch
import torch.distributed as dist
import torch.nn.functional as F


def gather_keeping_local_gradient(local, group=None):
    """[W * n, d] concatenation of every rank's `local` [n, d]; the local slice stays attached to the
    autograd graph, the others are constants (train_dense_retriever.py:131-153)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    detached = local.detach().contiguous()
    parts = [torch.empty_like(detached) for _ in range(world)]
    dist.all_gather(parts, detached, group=group)
    parts[rank] = local
    return torch.cat(parts, dim=0).contiguous()


def in_batch_labels(local_batch_size, local_context_size, world_size, train_with_neg, device=None):
    """Position of each question's positive context in the gathered context list (:163-175): with hard
    negatives every rank contributes `local_context_size` rows of which the first `local_batch_size`
    are positives."""
    if train_with_neg:
        labels = [j for r in range(world_size)
                  for j in range(r * local_context_size, r * local_context_size + local_batch_size)]
        return torch.tensor(labels, dtype=torch.int64, device=device)
    return torch.arange(world_size * local_batch_size, dtype=torch.int64, device=device)


def in_batch_negative_loss(query_logits, context_logits, hidden_size, retriever_score_scaling=True,
                           train_with_neg=False, group=None):
    """(loss, stats): loss already multiplied by the data-parallel world size like the reference's
    (:190), stats = {'lm loss': mean NLL, 'correct_prediction_count': hits} before any cross-rank
    averaging (the reference averages them for logging only)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    local_batch_size, local_context_size = query_logits.shape[0], context_logits.shape[0]
    if train_with_neg and local_context_size < local_batch_size:
        raise ValueError("with hard negatives the context batch holds the positives first, then the negatives")
    if not train_with_neg and local_context_size != local_batch_size:
        raise ValueError("without hard negatives every question brings exactly one context")
    all_q = gather_keeping_local_gradient(query_logits.float(), group)
    all_c = gather_keeping_local_gradient(context_logits.float(), group)
    scores = torch.matmul(all_q, all_c.transpose(0, 1))
    if retriever_score_scaling:
        scores = scores / math.sqrt(hidden_size)
    labels = in_batch_labels(local_batch_size, local_context_size, world, train_with_neg, device=scores.device)
    log_probs = F.log_softmax(scores, dim=1)
    nll = F.nll_loss(log_probs, labels, reduction="mean")
    correct = (log_probs.argmax(dim=1) == labels).sum().float()
    return nll * world, {"lm loss": nll.detach(), "correct_prediction_count": correct}


def forward_step(model, query_tokens, query_types, context_tokens, context_types, neg_context_tokens=None,
                 neg_context_types=None, hidden_size=768, retriever_score_scaling=True, group=None):
    """One training forward (:89-196) on a `DualEncoder`: embeds, exchanges, returns (loss, stats).
    The dense masks of the reference batch are not needed (the kernels derive them from the ids)."""
    train_with_neg = neg_context_tokens is not None
    if train_with_neg:
        context_tokens = torch.cat([context_tokens, neg_context_tokens])
        context_types = torch.cat([context_types, neg_context_types])
    query_logits, context_logits = model(query_tokens, None, query_types, context_tokens, None, context_types)
    return in_batch_negative_loss(query_logits, context_logits, hidden_size, retriever_score_scaling,
                                  train_with_neg, group)
