"""fp32 CPU restatement of the scoring / gather / loss / optimizer arithmetic of the reference task.

Follows /root/reference/dpr_scale/task/dpr_task.py:
  sim_score            :98-105   (q @ c.T, scores[mask] = -inf)
  in_batch_loss        :197, :209-212 (mask.repeat, /= temperature, CrossEntropyLoss mean)
  non_in_batch_mask    :199-207
  gather_for_rank      :163-195 (all_gather of detached reps; local slice re-inserted; label offsets)
  lr_lambda            :135-142
  rank_metrics         :235-246
and torch.optim.AdamW / clip_grad_norm_ as configured by conf/task/optim/adamw.yaml, conf/trainer/gpu_1_host.yaml:8.
"""
import math

import torch
import torch.nn.functional as F


def sim_score(q, c, mask=None):
    s = q @ c.T
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return s


def in_batch_loss(q, c, ctx_mask, pos_idx, temperature=1.0):
    """Returns (loss, logits) exactly as training_step computes them for the in_batch_negatives branch."""
    m = ctx_mask.unsqueeze(0).expand(q.shape[0], -1)
    logits = sim_score(q, c, m) / temperature
    return F.cross_entropy(logits, pos_idx), logits


def non_in_batch_mask(ctx_mask, pos_idx, num_q):
    per = int(ctx_mask.shape[0] / num_q)
    m = torch.ones(num_q, ctx_mask.shape[0], dtype=torch.bool)
    for i, p in enumerate(pos_idx.tolist()):
        m[i, p:p + per] = ctx_mask[p:p + per]
    return m


def gather_for_rank(rank, q_all, c_all, labels_all, mask_all):
    """q_all/c_all/labels_all/mask_all: lists (one entry per rank) as produced by the all_gather at :174-176.
    Entry `rank` of q_all/c_all is the grad-carrying local tensor; others are treated as constants."""
    offset = 0
    qs, cs, labs = [], [], []
    for i in range(len(q_all)):
        qs.append(q_all[i] if i == rank else q_all[i].detach())
        cs.append(c_all[i] if i == rank else c_all[i].detach())
        labs.append(labels_all[i] + offset)
        offset += c_all[i].shape[0]
    return torch.cat(qs), torch.cat(cs), torch.cat(labs), torch.cat(mask_all)


def lr_lambda(step, warmup_steps, training_steps):
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    return max(0.0, float(training_steps - step) / float(max(1, training_steps - warmup_steps)))


def rank_metrics(scores, labels, k=1):
    order = torch.argsort(scores, dim=1, descending=True, stable=True)
    rank = 0
    mrr = 0.0
    hit = 0
    for i, lab in enumerate(labels.tolist()):
        pos = int((order[i] == lab).nonzero()[0, 0])
        rank += pos + 1
        hit += int(pos < k)
        mrr += 1.0 / (pos + 1)
    return rank, mrr, hit


def clip_coef(grads, max_norm):
    total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    return min(1.0, max_norm / (total + 1e-6)), total


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One decoupled-weight-decay Adam update (torch.optim.AdamW semantics, amsgrad off), in place."""
    p.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p.addcdiv_(m, denom, value=-lr / bc1)
