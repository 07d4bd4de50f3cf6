#!/usr/bin/env python3
"""Writes tests/golden/retrieval_small.npz by running the UNMODIFIED reference retrieval script
(/root/reference/dpr_scale/run_retrieval_pytorch.py) in the authoring container:

  * `search_index` (:141-176) is imported and called as is, per index segment;
  * `main` (:192-300) runs end to end on small on-disk inputs (reps_*.pkl, query_reps.pkl, passage / question tables)
    with --shard 3, once in DPR-json and once in trec format; the run files it writes are stored verbatim.

The module imports `ujson`, `hydra` and `pytorch_lightning` (absent here) on its way to the table readers and hard-codes
`.cuda(0)`: the generator stubs the three imports exactly like make_golden_data.py does and patches `Tensor.cuda` to
the identity, so every arithmetic call (einsum('ik,jk->ij') on fp16 tensors, torch.topk, torch.gather) is the
reference's own line, executed by torch on the CPU.  Nothing of the reference is copied into this file.

The inputs are built so that, inside every query's top-(k+1), the fp16 scores are pairwise distinct: `torch.topk`'s order
among EQUAL scores is unspecified (and differs between its CPU and CUDA kernels), so only then is the id list a
property of the algorithm rather than of the backend.

Run from the repo root: python tests/golden/make_golden_retrieval.py
"""
import argparse
import json
import os
import pickle
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_stubs():
    sys.modules["ujson"] = json
    hydra = types.ModuleType("hydra")
    hu = types.ModuleType("hydra.utils")
    hu.instantiate = lambda conf, *a, **k: (_ for _ in ()).throw(RuntimeError("not used"))
    hydra.utils = hu
    sys.modules["hydra"], sys.modules["hydra.utils"] = hydra, hu
    pl = types.ModuleType("pytorch_lightning")

    class LightningDataModule:
        def __init__(self):
            self.trainer = None
    pl.LightningDataModule = LightningDataModule
    sys.modules["pytorch_lightning"] = pl
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # the script hard-codes .cuda(0)


def distinct_topk(q, shards, k):
    """True when every query's k+1 best fp16 scores over the concatenated index are pairwise distinct."""
    full = torch.cat(shards).to(torch.float16)
    s = torch.einsum("ik,jk->ij", (q.to(torch.float16), full)).float()
    top = torch.topk(s, k + 1, dim=-1).values
    return bool((top[:, :-1] != top[:, 1:]).all())


def main():
    install_stubs()
    import dpr_scale.run_retrieval_pytorch as ref
    d, nq, topk, batch = 64, 23, 10, 8
    shard_sizes = [700, 700, 700]
    seed = 20240917
    while True:
        g = torch.Generator().manual_seed(seed)
        q = torch.randn(nq, d, generator=g)
        shards = [torch.randn(n, d, generator=g) for n in shard_sizes]
        if distinct_topk(q, shards, topk):
            break
        seed += 1
    out = {"queries": q.numpy(), "topk": topk, "batch": batch, "seed": seed}
    for j, sh in enumerate(shards):
        index = sh.to(torch.float16)                         # build_index(): .to(torch.float16).cuda(0)
        s, i = ref.search_index(q, index, batch, topk)       # the reference's own function
        out[f"shard{j}"] = sh.numpy()
        out[f"scores{j}"] = np.asarray(s, dtype=np.float64)
        out[f"index{j}"] = np.asarray(i, dtype=np.float64)
    # ---- the reference's main(), end to end
    tmp = tempfile.mkdtemp(prefix="dprb_retr_")
    try:
        for j, sh in enumerate(shards):
            with open(os.path.join(tmp, f"reps_{j:04}.pkl"), "wb") as f:
                pickle.dump(sh, f, protocol=4)               # what GenerateEmbeddingsTask writes (dpr_eval_task.py:46)
        with open(os.path.join(tmp, "query_reps.pkl"), "wb") as f:
            pickle.dump(q, f, protocol=4)
        n_pass = sum(shard_sizes)
        ptsv = "id\ttext\ttitle\n" + "".join(f'p{i}\t"passage ""{i}"" text"\ttitle {i}\n' for i in range(n_pass))
        qcsv = "".join(f"question number {i}?\t['answer {i}', \"alt {i}\"]\n" for i in range(nq))
        qtsv = "".join(f"q{i}\tquestion number {i}?\n" for i in range(nq))
        for name, text in (("psgs.tsv", ptsv), ("q.csv", qcsv), ("q.tsv", qtsv)):
            with open(os.path.join(tmp, name), "w") as f:
                f.write(text)
        out["passages_tsv"], out["questions_csv"], out["questions_tsv"] = ptsv, qcsv, qtsv
        for fmt, qfile, trec in (("json", "q.csv", False), ("trec", "q.tsv", True)):
            runfile = os.path.join(tmp, f"run.{fmt}")
            argv = ["--ctx_embeddings_dir", tmp, "--query_emb_path", os.path.join(tmp, "query_reps.pkl"),
                    "--questions_tsv_path", os.path.join(tmp, qfile), "--passages_tsv_path", os.path.join(tmp, "psgs.tsv"),
                    "--output_runfile_path", runfile, "--topk", str(topk), "--batch", str(batch), "--shard", "3"]
            if trec:
                argv += ["--trec_format", "--run_name", "golden"]
            ref.main(ref.get_parser().parse_args(argv), ref.get_logger())
            out[f"run_{fmt}"] = open(runfile).read()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    # merged lists, parsed back from the trec run (scores are printed with repr precision)
    ms, mi = np.zeros((nq, topk)), np.zeros((nq, topk))
    for line in out["run_trec"].splitlines():
        qid, _, pid, rank, score, _ = line.split()
        ms[int(qid[1:]), int(rank) - 1] = float(score)
        mi[int(qid[1:]), int(rank) - 1] = int(pid[1:])
    out["merged_scores"], out["merged_index"] = ms, mi
    np.savez_compressed(os.path.join(HERE, "retrieval_small.npz"), **out)
    print("wrote retrieval_small.npz (seed %d)" % seed, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
