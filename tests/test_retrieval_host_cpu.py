"""Host-side logic of dpr_scale_b200/run_retrieval.py (file readers, run-file writer) - no GPU needed.
Formats follow /root/reference/dpr_scale/datamodule/dpr.py:80-159 and run_retrieval_pytorch.py:96-137, :232-300."""
import json

import numpy as np

from dpr_scale_b200 import run_retrieval as RR


def _write_inputs(tmp_path, n_pass=12):
    ptsv = tmp_path / "psgs.tsv"
    with open(ptsv, "w") as f:
        f.write("id\ttext\ttitle\n")
        for i in range(n_pass):
            f.write(f'{i + 1}\t"passage ""{i}"" text"\ttitle {i}\n')
    qcsv = tmp_path / "q.csv"
    with open(qcsv, "w") as f:
        f.write("who wrote x?\t['a', \"b c\"]\n")
        f.write('"what is ""y""?"\t[\'d\']\n')
    qtsv = tmp_path / "q.tsv"
    with open(qtsv, "w") as f:
        f.write("q7\twho wrote x?\n3\twhat is y?\n")
    return ptsv, qcsv, qtsv


def test_table_readers(tmp_path):
    ptsv, qcsv, qtsv = _write_inputs(tmp_path)
    p = RR.Passages(str(ptsv))
    assert len(p) == 12 and p.columns == ["id", "text", "title"]
    assert p[0] == {"id": "1", "text": 'passage "0" text', "title": "title 0"}
    assert p[np.float64(11.0)]["id"] == "12"            # the reference indexes with float row ids
    q = list(RR.Questions(str(qcsv), False))
    assert q == [{"question": "who wrote x?", "answers": ["a", "b c"]}, {"question": 'what is "y"?', "answers": ["d"]}]
    t = list(RR.Questions(str(qtsv), True))
    assert t == [{"id": "q7", "question": "who wrote x?"}, {"id": "3", "question": "what is y?"}]


def test_write_run_json_and_trec(tmp_path):
    ptsv, qcsv, qtsv = _write_inputs(tmp_path)
    p = RR.Passages(str(ptsv))
    scores = np.array([[3.5, 2.25, 1.0], [9.0, 8.0, 7.0]])
    idx = np.array([[4, 0, 11], [2, 3, 5]])
    out = tmp_path / "o" / "run.json"
    RR.write_run(str(out), p, list(RR.Questions(str(qcsv), False)), scores, idx, False)
    d = json.load(open(out))
    assert [c["id"] for c in d[0]["ctxs"]] == ["5", "1", "12"] and d[0]["ctxs"][1]["score"] == 2.25
    assert d[1]["question"] == 'what is "y"?' and d[1]["answers"] == ["d"] and d[1]["id"] == 1
    assert d[0]["ctxs"][0]["text"] == 'passage "4" text' and d[0]["ctxs"][0]["title"] == "title 4"
    trec = tmp_path / "run.trec"
    RR.write_run(str(trec), p, list(RR.Questions(str(qtsv), True)), scores, idx, True, run_name="b200",
                 ignore_identical_ids=True)
    lines = open(trec).read().splitlines()
    # query "3" retrieved passage id "3" at rank 1: dropped by --ignore_identical_ids, ranks keep their numbers
    assert lines == ["q7 Q0 5 1 3.5 b200", "q7 Q0 1 2 2.25 b200", "q7 Q0 12 3 1.0 b200",
                     "3 Q0 4 2 8.0 b200", "3 Q0 6 3 7.0 b200"]


def test_parser_matches_reference_flags():
    a = RR.get_parser().parse_args([])
    assert (a.topk, a.batch, a.shard, a.run_name, a.trec_format, a.ignore_identical_ids) == (100, 100, 1, "dpr", False, False)
    for flag in ("ctx_embeddings_dir", "query_emb_path", "questions_tsv_path", "passages_tsv_path", "output_runfile_path"):
        assert getattr(a, flag) == ""


def test_main_flow_with_emulated_search(tmp_path, monkeypatch):
    """Whole script flow (reps_* pickles -> segments -> merge -> run file) on the CPU with the two kernels replaced by
    torch emulations - host logic only; the kernels themselves are checked on the GPU (tests/test_retrieval_gpu.py)."""
    import pickle

    import torch

    from dpr_scale_b200 import ops

    def fake_search(q, c, k, index_offset=0, reference_ranking=False):
        s = q.float() @ c.float().T
        v, i = torch.sort(s, dim=1, descending=True, stable=True)
        return v[:, :k].contiguous(), i[:, :k] + index_offset

    def fake_merge(s, idx, k):
        v, o = torch.sort(s, dim=1, descending=True, stable=True)
        return v[:, :k].contiguous(), torch.gather(idx, 1, o[:, :k])
    monkeypatch.setattr(ops, "search_topk", fake_search)
    monkeypatch.setattr(ops, "topk_merge", fake_merge)
    g = torch.Generator().manual_seed(0)
    emb = tmp_path / "emb"
    emb.mkdir()
    shards = [torch.randn(40, 16, generator=g) for _ in range(4)]
    for r, t in enumerate(shards):
        pickle.dump(t, open(emb / f"reps_{r:04}.pkl", "wb"), protocol=4)
    q = torch.randn(3, 16, generator=g)
    pickle.dump(q, open(emb / "query_reps.pkl", "wb"), protocol=4)
    with open(tmp_path / "p.tsv", "w") as f:
        f.write("id\ttext\ttitle\n" + "".join(f"p{i}\ttext {i}\ttitle {i}\n" for i in range(160)))
    with open(tmp_path / "q.tsv", "w") as f:
        f.write("".join(f"q{i}\tquestion {i}\n" for i in range(3)))
    full = torch.cat(shards).half().float()
    want = torch.sort(q.half().float() @ full.T, dim=1, descending=True, stable=True)[1][:, :5]
    for shard in (1, 2, 4):
        out = tmp_path / f"run{shard}.trec"
        RR.main(RR.get_parser().parse_args([
            "--ctx_embeddings_dir", str(emb), "--questions_tsv_path", str(tmp_path / "q.tsv"), "--passages_tsv_path",
            str(tmp_path / "p.tsv"), "--output_runfile_path", str(out), "--topk", "5", "--shard", str(shard),
            "--trec_format", "--device", "cpu"]))
        rows = [l.split() for l in open(out).read().splitlines()]
        assert [r[2] for r in rows] == [f"p{int(j)}" for j in want.flatten()], shard
        assert [int(r[3]) for r in rows] == [1, 2, 3, 4, 5] * 3
