"""End-to-end retrieval on the GPU: reps_* pickles -> dpr_scale_b200.run_retrieval.main -> run file, against the oracle
restatement of run_retrieval_pytorch.py (oracle/retrieval.py), plus a full-size property test."""
import json
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(tmp_path, n_files=4, rows=700, d=64, nq=9, seed=3):
    g = torch.Generator().manual_seed(seed)
    emb = tmp_path / "emb"
    emb.mkdir()
    shards = []
    for r in range(n_files):
        t = torch.randn(rows, d, generator=g)
        shards.append(t)
        with open(emb / f"reps_{r:04}.pkl", "wb") as f:
            pickle.dump(t, f, protocol=4)
    q = torch.randn(nq, d, generator=g)
    with open(emb / "query_reps.pkl", "wb") as f:
        pickle.dump(q, f, protocol=4)
    with open(tmp_path / "psgs.tsv", "w") as f:
        f.write("id\ttext\ttitle\n")
        for i in range(n_files * rows):
            f.write(f"p{i}\ttext {i}\ttitle {i}\n")
    with open(tmp_path / "q.tsv", "w") as f:
        for i in range(nq):
            f.write(f"q{i}\tquestion {i}\n")
    return emb, q, shards


@pytest.mark.parametrize("shard", [1, 2])
def test_run_retrieval_trec_matches_oracle(tmp_path, shard):
    from dpr_scale_b200 import run_retrieval as RR
    from oracle import retrieval as R
    emb, q, shards = _setup(tmp_path)
    out = tmp_path / "run.trec"
    args = RR.get_parser().parse_args([
        "--ctx_embeddings_dir", str(emb), "--questions_tsv_path", str(tmp_path / "q.tsv"), "--passages_tsv_path",
        str(tmp_path / "psgs.tsv"), "--output_runfile_path", str(out), "--topk", "10", "--shard", str(shard),
        "--trec_format"])
    RR.main(args)
    full = np.concatenate([s.numpy() for s in shards], axis=0)
    exact = R.exact_scores(q.numpy(), full)
    es, ei = R.topk_desc(exact, 11)
    lines = [l.split() for l in open(out).read().splitlines()]
    assert len(lines) == 9 * 10
    for n, (qid, q0, pid, rank, score, name) in enumerate(lines):
        r, j = divmod(n, 10)
        assert qid == f"q{r}" and q0 == "Q0" and int(rank) == j + 1 and name == "dpr"
        row = int(pid[1:])
        # score written = fp16 rounding of the fp32-accumulated product (the reference's scores are fp16 einsum outputs)
        assert abs(float(score) - exact[r, row]) <= np.spacing(np.float16(abs(exact[r, row]))) * 0.51 + 1e-3
        gap_lo = es[r, j] - es[r, j + 1]
        gap_hi = es[r, j - 1] - es[r, j] if j else np.inf
        if min(gap_lo, gap_hi) > 2e-3:
            assert row == ei[r, j], (r, j, row, ei[r, j])


def test_run_retrieval_json_output(tmp_path):
    from dpr_scale_b200 import run_retrieval as RR
    emb, q, shards = _setup(tmp_path, n_files=1, rows=300, nq=3)
    with open(tmp_path / "q.csv", "w") as f:
        for i in range(3):
            f.write(f"question {i}\t['ans {i}']\n")
    out = tmp_path / "run.json"
    RR.main(RR.get_parser().parse_args([
        "--ctx_embeddings_dir", str(emb), "--questions_tsv_path", str(tmp_path / "q.csv"), "--passages_tsv_path",
        str(tmp_path / "psgs.tsv"), "--output_runfile_path", str(out), "--topk", "5"]))
    d = json.load(open(out))
    assert len(d) == 3 and all(len(e["ctxs"]) == 5 for e in d)
    assert d[2]["answers"] == ["ans 2"] and d[2]["id"] == 2
    s = [c["score"] for c in d[0]["ctxs"]]
    assert s == sorted(s, reverse=True)
    best = int(np.argmax(q[0].half().float().numpy() @ shards[0].half().float().numpy().T))
    assert d[0]["ctxs"][0]["id"] == f"p{best}" and d[0]["ctxs"][0]["title"] == f"title {best}"


def test_search_full_size_properties():
    """2 M x 768 fp16 index (3 GB), 128 queries, k = 100: properties that need no CPU oracle.
    (1) planted rows: each query's own vector (scaled) is inserted at a known row and must come back at rank 1;
    (2) the k-th returned score bounds every row NOT returned (checked against a chunked fp32 torch matmul);
    (3) searching the two halves separately and merging equals the single search (scores exactly, ids exactly)."""
    from dpr_scale_b200 import ops
    dev = "cuda"
    N, d, Q, k = 2_000_000, 768, 128, 100
    g = torch.Generator(device=dev).manual_seed(11)
    corpus = torch.empty(N, d, dtype=torch.float16, device=dev)
    for s in range(0, N, 1 << 19):
        e = min(N, s + (1 << 19))
        corpus[s:e] = torch.randn(e - s, d, generator=g, device=dev).to(torch.float16)
    q = torch.randn(Q, d, generator=g, device=dev).to(torch.float16)
    planted = torch.randint(0, N, (Q,), generator=g, device=dev)
    planted = torch.unique(planted)
    nq = planted.numel()
    corpus[planted] = (q[:nq].float() * 2.0).to(torch.float16)
    s, i = ops.search_topk(q, corpus, k)
    assert torch.equal(i[:nq, 0], planted)
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    kth = s[:, -1]
    count_above = torch.zeros(Q, dtype=torch.int64, device=dev)
    for c0 in range(0, N, 1 << 18):
        sc = q.float() @ corpus[c0:c0 + (1 << 18)].float().T
        count_above += (sc > kth[:, None] + 1e-2).sum(1)
    assert bool((count_above <= k - 1).all()), "rows better than the k-th result were missed"
    h = N // 2 + 77
    s1, i1 = ops.search_topk(q, corpus[:h], k)
    s2, i2 = ops.search_topk(q, corpus[h:], k, index_offset=h)
    ms, mi = ops.topk_merge(torch.cat([s1, s2], 1).contiguous(), torch.cat([i1, i2], 1).contiguous(), k)
    assert torch.equal(ms, s) and torch.equal(mi, i)


def _dist_worker(rank, world, port, root, ret):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import glob
    from dpr_scale_b200 import run_retrieval as RR
    paths = sorted(glob.glob(os.path.join(root, "emb", "reps_*")))
    with open(os.path.join(root, "emb", "query_reps.pkl"), "rb") as f:
        q = pickle.load(f)
    s, i = RR.search_distributed(q, paths, 1, 100, 10, device=f"cuda:{rank}")
    s1, i1 = RR.search_segments(q, paths, 1, 100, 10, device=f"cuda:{rank}")        # whole index on one GPU
    ret[rank] = (bool(torch.equal(s, s1)), bool(torch.equal(i, i1)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_rank_sharded_search_equals_single_gpu(tmp_path):
    """Index sharded over 2 ranks (NCCL all-gather of the per-rank lists + merge) == one GPU holding the whole index,
    bit for bit (same fp32 scores, same tie order)."""
    import torch.multiprocessing as mp
    _setup(tmp_path, n_files=4, rows=900, d=128, nq=17, seed=5)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dist_worker, args=(2, 29677, str(tmp_path), ret), nprocs=2, join=True)
    assert ret[0] == (True, True) and ret[1] == (True, True)


def test_reference_ranking_reproduces_the_reference_run_files(tmp_path):
    """`--reference_ranking` (rank by the fp16-rounded score): dpr_scale_b200.run_retrieval.main on the golden inputs
    writes the SAME run files, byte for byte, that the unmodified reference main() wrote
    (tests/golden/make_golden_retrieval.py; its inputs have pairwise distinct fp16 scores inside every top-(k+1))."""
    import os
    from dpr_scale_b200 import run_retrieval as RR
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "retrieval_small.npz"))
    emb = tmp_path / "emb"
    emb.mkdir()
    for j in range(3):
        with open(emb / f"reps_{j:04}.pkl", "wb") as f:
            pickle.dump(torch.from_numpy(g[f"shard{j}"]), f, protocol=4)
    with open(emb / "query_reps.pkl", "wb") as f:
        pickle.dump(torch.from_numpy(g["queries"]), f, protocol=4)
    for name in ("passages_tsv", "questions_csv", "questions_tsv"):
        with open(tmp_path / name, "w") as f:
            f.write(str(g[name]))
    k = int(g["topk"])
    for fmt, qfile, extra in (("json", "questions_csv", []), ("trec", "questions_tsv", ["--trec_format", "--run_name", "golden"])):
        out = tmp_path / f"run.{fmt}"
        args = RR.get_parser().parse_args(
            ["--ctx_embeddings_dir", str(emb), "--questions_tsv_path", str(tmp_path / qfile),
             "--passages_tsv_path", str(tmp_path / "passages_tsv"), "--output_runfile_path", str(out),
             "--topk", str(k), "--shard", "3", "--reference_ranking"] + extra)
        RR.main(args)
        assert open(out).read() == str(g[f"run_{fmt}"]), fmt
    # per segment, against the reference's own search_index outputs
    for j in range(3):
        idx = torch.from_numpy(g[f"shard{j}"]).to("cuda", torch.float16)
        s, i = RR.search_index(torch.from_numpy(g["queries"]), idx, 8, k, reference_ranking=True)
        sc, ids, gold = s.cpu().numpy().astype(np.float64), i.cpu().numpy(), g[f"index{j}"].astype(np.int64)
        assert np.array_equal(sc, g[f"scores{j}"])
        # inside ONE segment fp16 scores may tie (the generator only separates the merged top-(k+1)); torch.topk's order
        # among equal scores is unspecified, so ids are compared where the score is unique in its row
        uniq = np.ones_like(sc, dtype=bool)
        uniq[:, 1:] &= sc[:, 1:] != sc[:, :-1]
        uniq[:, :-1] &= sc[:, :-1] != sc[:, 1:]
        uniq[:, -1] = False
        assert uniq.mean() > 0.8 and np.array_equal(ids[uniq], gold[uniq])
