#!/usr/bin/env python3
"""Writes tests/golden/retrieval_small.npz: inputs and the outputs of the reference's retrieval arithmetic
(/root/reference/dpr_scale/run_retrieval_pytorch.py:141-176, :210-230, :272-277) executed by torch on the CPU -
the same calls (einsum('ik,jk->ij') on fp16 tensors, torch.topk, torch.gather) minus `.cuda(0)`.

Run from the repo root: python tests/golden/make_golden_retrieval.py
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_search_index(query_embs, corpus_embs, batch, topk):
    n = query_embs.shape[0]
    all_scores = np.zeros((n, topk))
    all_results = np.zeros((n, topk))
    if batch > n:
        scores = torch.einsum("ik,jk->ij", (query_embs.to(torch.float16), corpus_embs))
        s, c = torch.topk(scores, dim=-1, k=topk)
        return s.float().numpy().astype(np.float64), c.numpy().astype(np.float64)
    i = -1
    for i in range(n // batch):
        scores = torch.einsum("ik,jk->ij", (query_embs[i * batch:(i + 1) * batch].to(torch.float16), corpus_embs))
        s, c = torch.topk(scores, dim=-1, k=topk)
        all_results[i * batch:(i + 1) * batch] = c
        all_scores[i * batch:(i + 1) * batch] = s.float()
    if (i + 1) * batch < n:
        scores = torch.einsum("ik,jk->ij", (query_embs[(i + 1) * batch:].to(torch.float16), corpus_embs))
        s, c = torch.topk(scores, dim=-1, k=topk)
        all_scores[(i + 1) * batch:] = s.float()
        all_results[(i + 1) * batch:] = c
    return all_scores, all_results


def main():
    g = torch.Generator().manual_seed(20240917)
    d, nq, topk, batch = 64, 23, 10, 8
    shard_sizes = [700, 700, 700]
    q = torch.randn(nq, d, generator=g)
    shards = [torch.randn(n, d, generator=g) for n in shard_sizes]
    out = {"queries": q.numpy(), "topk": topk, "batch": batch}
    all_s, all_i, offset = [], [], 0
    for j, sh in enumerate(shards):
        index = sh.to(torch.float16)                     # build_index(): .to(torch.float16)
        s, i = ref_search_index(q, index, batch, topk)
        out[f"shard{j}"] = sh.numpy()
        out[f"scores{j}"] = s
        out[f"index{j}"] = i
        all_s.append(s)
        all_i.append(i + offset)
        offset += len(index)
    all_s = np.concatenate(all_s, axis=1)
    all_i = np.concatenate(all_i, axis=1)
    ms, idx = torch.topk(torch.tensor(all_s), dim=-1, k=topk)      # :274
    mi = torch.gather(torch.tensor(all_i), 1, idx)                 # :275
    out["merged_scores"] = ms.numpy()
    out["merged_index"] = mi.numpy()
    np.savez_compressed(os.path.join(HERE, "retrieval_small.npz"), **out)
    print("wrote retrieval_small.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
