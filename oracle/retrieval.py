"""CPU restatement of the reference's brute-force retrieval (TEST INFRASTRUCTURE ONLY - never imported by the
product path).

Follows /root/reference/dpr_scale/run_retrieval_pytorch.py:
  search_index   :141-176  scores = einsum('ik,jk->ij', q.half(), corpus.half()) (fp16 result); torch.topk(k);
                           queries processed in blocks of `batch` (the block loop changes nothing numerically)
  build_index    :178-190  concatenation of the reps_* shards, cast to fp16
  shard merge    :210-230  per index shard: topk, row ids + running offset, results concatenated along dim 1
                 :272-277  final torch.topk over the concatenation + gather of the ids

Pinned against torch CPU executions of exactly those calls (tests/golden/retrieval_small.npz, written by
tests/golden/make_golden_retrieval.py); the reference module itself cannot be imported here (ujson / hydra are
absent and it hard-codes .cuda(0)).  torch.topk leaves the order of equal scores unspecified; this restatement
orders ties by ascending row id, and the parity tests compare score vectors, and ids only where scores are distinct.
"""
import numpy as np


def _as_fp16(x):
    return np.asarray(x).astype(np.float16)


def exact_scores(query_embs, corpus_embs):
    """float64 inner products of the fp16-rounded operands (what every fp16/fp32-accumulating GEMM approximates)."""
    return _as_fp16(query_embs).astype(np.float64) @ _as_fp16(corpus_embs).astype(np.float64).T


def topk_desc(scores, k):
    """Top-k per row, descending score, ties by ascending column."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, axis=1), order


def search_index(query_embs, corpus_embs, batch, topk, fp16_scores=True):
    """run_retrieval_pytorch.py:141-176.  Returns (all_scores [Q, k], all_results [Q, k])."""
    q = _as_fp16(query_embs)
    c = _as_fp16(corpus_embs)
    n = q.shape[0]
    all_scores = np.zeros((n, topk))
    all_results = np.zeros((n, topk), dtype=np.int64)
    starts = [0] if batch > n else list(range(0, n, batch))
    for s in starts:
        e = n if batch > n else min(n, s + batch)
        scores = q[s:e].astype(np.float32) @ c.astype(np.float32).T          # fp32 accumulate, as cuBLAS / mkldnn do
        if fp16_scores:
            scores = scores.astype(np.float16).astype(np.float32)           # einsum output dtype is fp16
        sv, si = topk_desc(scores, topk)
        all_scores[s:e] = sv
        all_results[s:e] = si
    return all_scores, all_results


def build_index(shards):
    """run_retrieval_pytorch.py:178-190 (without the .cuda)."""
    return np.concatenate([np.asarray(s, dtype=np.float32) for s in shards], axis=0).astype(np.float16)


def search_shards(query_embs, shards, batch, topk, fp16_scores=True):
    """run_retrieval_pytorch.py:204-230 + :272-277: search every index shard, offset its row ids, merge."""
    all_scores, all_indexes, offset = [], [], 0
    for shard in shards:
        s, i = search_index(query_embs, shard, batch, topk, fp16_scores)
        all_scores.append(s)
        all_indexes.append(i + offset)
        offset += len(shard)
    all_scores = np.concatenate(all_scores, axis=1)
    all_indexes = np.concatenate(all_indexes, axis=1)
    if len(shards) == 1:
        return all_scores, all_indexes
    sv, order = topk_desc(all_scores, topk)
    return sv, np.take_along_axis(all_indexes, order, axis=1)
