"""Reader / retriever losses of the EMDR2 step (forward), on the fused log-prob kernel.

Mirrors reference tasks/openqa/e2eqa/train_e2eqa.py:72-123 (`get_loss_and_retriever_utility`),
:184-214 (`get_kl_div_retriever`) and the reader cross-entropy of `_cross_entropy_forward_step`
(:152-160): same names, arguments and return values.  The reference materialises
log_softmax over [B, K, L, V] in fp32 (1.57 GB at B=8, K=50, L=32, V=30720) and gathers the label
column; here `ops.token_logprob` reads the 16-bit logits once and emits the [B, K, L] gold
log-probabilities directly; everything after that is [B, K, L]-sized torch arithmetic.
`*_from_gold` variants take the gathered log-probabilities so the CPU tests can exercise the
post-gather arithmetic without a GPU.
"""
import torch
import torch.nn.functional as F

from . import autograd as ag
from . import ops


def gold_log_probs(lm_logits, labels, loss_mask):
    """[B, K, L] log p(label | ...) from logits [B, K, L, V]; masked labels are read as id 0 like
    the reference's `labels.masked_fill(~loss_mask, 0)` (:88)."""
    labels = labels.masked_fill(~loss_mask.to(torch.bool), 0)
    topk = lm_logits.shape[1]
    tiled = labels.unsqueeze(1).expand(-1, topk, -1)
    lp, _ = ops.token_logprob(lm_logits, tiled.contiguous())
    return lp, labels


def loss_and_retriever_utility_from_gold(gold, topk_log_probs, labels, loss_mask, eos_id):
    """Arithmetic of :99-123 on gathered gold log-probs [B, K, L] (labels already mask-filled)."""
    topk_log_probs = topk_log_probs.float()
    joint = topk_log_probs.unsqueeze(-1) + gold
    marginal = torch.logsumexp(joint, dim=1)
    lm_loss = -1 * torch.sum(marginal * loss_mask) / torch.sum(loss_mask)
    utility = marginal - gold[:, -1, :]
    utility_mask = loss_mask.masked_fill(labels >= eos_id, 0)
    # the reference asserts this on the host (:116), which stalls the launching thread until the whole forward has run
    # and leaves the backward pass to be enqueued against an idle GPU; on a CUDA tensor the same condition is checked by
    # a device-side assertion instead (it fails just as loudly, at the next synchronisation)
    nonempty = torch.sum(utility_mask) > 0
    if nonempty.is_cuda:
        torch._assert_async(nonempty, "retriever-utility mask is empty")
    elif not nonempty:
        raise AssertionError("retriever-utility mask is empty")
    utility = torch.sum(utility * utility_mask) / torch.sum(utility_mask)
    null_block_lm_loss = -1 * torch.sum(gold[:, -1, :] * loss_mask) / torch.sum(loss_mask)
    return lm_loss, utility, null_block_lm_loss


def get_loss_and_retriever_utility(lm_logits, topk_log_probs, labels, loss_mask, eos_id):
    gold, labels = gold_log_probs(lm_logits, labels, loss_mask)
    return loss_and_retriever_utility_from_gold(gold, topk_log_probs, labels, loss_mask, eos_id)


def kl_div_retriever_from_gold(gold, topk_log_probs, loss_mask):
    teacher_log_probs = torch.sum(gold * loss_mask.unsqueeze(1), dim=2) / torch.sum(loss_mask.unsqueeze(1), dim=2)
    teacher_probs = torch.softmax(teacher_log_probs, dim=1)
    return F.kl_div(topk_log_probs.float(), teacher_probs, reduction='batchmean')


def get_kl_div_retriever(lm_logits, topk_log_probs, labels, loss_mask):
    gold, _ = gold_log_probs(lm_logits, labels, loss_mask)
    return kl_div_retriever_from_gold(gold, topk_log_probs, loss_mask)


def reader_cross_entropy(lm_logits, labels, loss_mask):
    """CrossEntropyLoss(reduction='none', ignore_index=0) summed under loss_mask (:156-160)."""
    lp = ag.token_logprob(lm_logits, labels.clamp(min=0))       # differentiable w.r.t. the logits
    loss_ = torch.where(labels == 0, torch.zeros_like(lp), -lp)
    return torch.sum(loss_.reshape(-1) * loss_mask.reshape(-1)) / loss_mask.sum()
