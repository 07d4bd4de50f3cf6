"""Input pipeline (dpr_scale_b200/datamodule/dpr.py, transforms/) vs golden batches produced by the UNMODIFIED reference
data code (tests/golden/make_golden_data.py -> data_batches.npz): every batch of every stage, bit for bit, through the
synchronous loader and through the background-thread loader; line index; distributed sampler orders."""
import os

import numpy as np
import pytest
import torch

from dpr_scale_b200.datamodule.dpr import (BatchStream, DenseRetrieverJsonlDataModule, LineFile,
                                           contiguous_shard_indices)
from dpr_scale_b200.transforms.hf_transform import BertTransform, HFTransform

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "data")
JSONL = os.path.join(DATA, "synth.jsonl")
CASES = {
    "a": dict(batch_size=4, num_negative=3, neg_ctx_sample=True, pos_ctx_sample=False, num_val_negative=2,
              num_test_negative=5, use_title=True),
    "b": dict(batch_size=5, num_negative=7, neg_ctx_sample=True, pos_ctx_sample=True, num_val_negative=7,
              num_test_negative=0, use_title=False, rel_sample=True),
    "c": dict(batch_size=15, num_negative=0, neg_ctx_sample=False, num_val_negative=1, use_title=False),
}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "data_batches.npz"))


@pytest.fixture(scope="module")
def model_dir(tmp_path_factory):
    from transformers import BertConfig
    d = tmp_path_factory.mktemp("tok")
    vocab = open(os.path.join(DATA, "vocab.txt")).read()
    BertConfig(vocab_size=len(vocab.split()), hidden_size=16, num_hidden_layers=1, num_attention_heads=1,
               intermediate_size=16).save_pretrained(d)
    (d / "vocab.txt").write_text(vocab)
    return str(d)


def _same(batch, gold, prefix):
    for key in ("query_ids", "contexts_ids"):
        names = [k.split("/")[-1] for k in gold.files if k.startswith(f"{prefix}/{key}/")]
        assert sorted(names) == sorted(batch[key].keys()), (names, list(batch[key].keys()))
        for kk in names:
            got = batch[key][kk]
            assert got.dtype == torch.int64
            assert np.array_equal(got.numpy(), gold[f"{prefix}/{key}/{kk}"]), f"{prefix}/{key}/{kk}"
    assert batch["pos_ctx_indices"].dtype == torch.int64 and batch["scores"].dtype == torch.float32
    assert batch["ctx_mask"].dtype == torch.bool
    for key in ("pos_ctx_indices", "scores", "ctx_mask"):
        assert np.array_equal(batch[key].numpy(), gold[f"{prefix}/{key}"]), f"{prefix}/{key}"


@pytest.mark.parametrize("prefetch,fast", [(0, False), (0, True), (3, True)])
@pytest.mark.parametrize("case", sorted(CASES))
def test_batches_equal_reference(gold, model_dir, case, prefetch, fast):
    tf = HFTransform(model_path=model_dir, max_seq_len=24)
    dm = DenseRetrieverJsonlDataModule(transform=tf, train_path=JSONL, val_path=JSONL, test_path=JSONL,
                                       prefetch_batches=prefetch, device_prefetch=False, fast_tokenize=fast,
                                       **CASES[case])
    np.random.seed(1234)
    for stage, loader in (("train", dm.train_dataloader()), ("valid", dm.val_dataloader()),
                          ("test", dm.test_dataloader())):
        want = int(gold[f"{case}/{stage}/num_batches"])
        assert len(loader) == want
        n = 0
        for i, batch in enumerate(loader):
            _same(batch, gold, f"{case}/{stage}/{i}")
            n += 1
        assert n == want


def test_bert_transform_and_pair_inputs(model_dir):
    t = BertTransform(model_path=model_dir, max_seq_len=8)
    out = t(["alpha bravo", "charlie delta echo foxtrot golf hotel india juliet"])
    assert out["input_ids"].shape == (2, 8) and out["attention_mask"][0].sum() == 4     # [CLS] a b [SEP]
    h = HFTransform(model_path=model_dir, max_seq_len=16, return_tensors=False)
    pair = h(["alpha"], ["bravo charlie"])
    assert pair["token_type_ids"][0] == [0, 0, 0, 1, 1, 1]


def test_encode_fast_equals_wrapper_call(model_dir):
    words = open(os.path.join(DATA, "vocab.txt")).read().split()[5:]
    rnd = np.random.RandomState(0)
    texts = [" ".join(rnd.choice(words, rnd.randint(0, 40))) for _ in range(200)] + ["", "zzz unknown-word !!"]
    for max_len in (8, 32):
        t = HFTransform(model_path=model_dir, max_seq_len=max_len)
        want, got = t(texts), t.encode_fast(texts)
        assert list(got.keys()) == list(want.keys())
        for k in want.keys():
            assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k
    t = HFTransform(model_path=model_dir, max_seq_len=16, add_special_tokens=False)
    assert torch.equal(t(texts[:50])["input_ids"], t.encode_fast(texts[:50])["input_ids"])


def test_line_index(gold, tmp_path):
    ds = LineFile(JSONL)
    assert len(ds) == int(gold["n_rows"])
    assert np.array_equal(np.array([len(ds[i]) for i in range(len(ds))]), gold["lines"])
    raw = open(JSONL, "rb").read().split(b"\n")
    assert ds[0] == raw[0] + b"\n" and ds[len(ds) - 1] == raw[-1]          # last line has no newline
    p = tmp_path / "h.tsv"
    p.write_text("id\ttext\n1\tx\n2\ty\n")
    h = LineFile(str(p), header=True)
    assert len(h) == 2 and h[0] == b"1\tx\n" and h[1] == b"2\ty\n"
    e = tmp_path / "empty.jsonl"
    e.write_text("")
    assert len(LineFile(str(e))) == 0
    with pytest.raises(KeyError):
        ds[len(ds)]


def test_sampler_orders_equal_reference(gold):
    n = int(gold["n_rows"])
    keys = [k for k in gold.files if k.startswith("sampler/")]
    assert len(keys) == 2 * (2 + 4 + 8 + 3)
    for k in keys:
        _, world, per_node, epoch, rank = k.split("/")
        got = contiguous_shard_indices(n, int(world), int(rank), int(per_node), True, 0, int(epoch), False)
        assert got == gold[k].tolist(), k


def test_stream_propagates_errors_and_stops_early():
    def bad(rows):
        if rows[0] == 4:
            raise ValueError("boom")
        return {"x": torch.tensor(rows)}
    s = BatchStream(list(range(10)), lambda: list(range(10)), 2, bad, prefetch_batches=2)
    it = iter(s)
    assert next(it)["x"].tolist() == [0, 1] and next(it)["x"].tolist() == [2, 3]
    with pytest.raises(ValueError):
        next(it)
    ok = BatchStream(list(range(100)), lambda: list(range(100)), 10, lambda r: {"x": torch.tensor(r)},
                     prefetch_batches=2, drop_last=True)
    for i, b in enumerate(ok):
        if i == 1:
            break                        # abandoning the iterator must stop the worker thread
    assert len(ok) == 10


# ------------------------------------------------------------------ embedding-generation datamodules (generate*.yaml)
@pytest.mark.parametrize("prefetch,fast", [(0, False), (2, True)])
def test_passages_datamodule_equals_reference(gold, model_dir, prefetch, fast):
    from dpr_scale_b200.datamodule.dpr import DenseRetrieverPassagesDataModule
    tf = HFTransform(model_path=model_dir, max_seq_len=24)
    path = os.path.join(DATA, "passages.tsv")
    for name, kw in (("t", dict(test_batch_size=5, use_title=True)), ("n", dict(test_batch_size=12))):
        dm = DenseRetrieverPassagesDataModule(tf, path, prefetch_batches=prefetch, device_prefetch=False,
                                              fast_tokenize=fast, **kw)
        n = 0
        for i, batch in enumerate(dm.test_dataloader()):
            for kk, v in batch["contexts_ids"].items():
                assert np.array_equal(v.numpy(), gold[f"passages/{name}/{i}/{kk}"]), (name, i, kk)
            assert batch["corpus_ids"] == gold[f"passages/{name}/{i}/corpus_ids"].tolist()
            n += 1
        assert n == int(gold[f"passages/{name}/num_batches"]) == len(dm.val_dataloader()) == len(dm.train_dataloader())


def test_queries_datamodule_equals_reference(gold, model_dir):
    from dpr_scale_b200.datamodule.dpr import DenseRetrieverQueriesDataModule
    tf = HFTransform(model_path=model_dir, max_seq_len=24)
    for name, fn, trec in (("csv", "questions.csv", False), ("tsv", "questions.tsv", True)):
        dm = DenseRetrieverQueriesDataModule(tf, os.path.join(DATA, fn), test_batch_size=3, trec_format=trec,
                                             device_prefetch=False)
        n = 0
        for i, batch in enumerate(dm.test_dataloader()):
            assert list(batch.keys()) == ["query_ids"]
            for kk, v in batch["query_ids"].items():
                assert np.array_equal(v.numpy(), gold[f"queries/{name}/{i}/{kk}"]), (name, i, kk)
            n += 1
        assert n == int(gold[f"queries/{name}/num_batches"])


def test_csv_dataset_and_test_sampler(gold):
    from dpr_scale_b200.datamodule.dpr import CSVDataset, QueryCSVDataset, contiguous_test_shard
    csv = CSVDataset(os.path.join(DATA, "passages.tsv"))
    assert len(csv) == int(gold["csv_len"]) and csv.columns == ["id", "text", "title"]
    assert csv[3]["text"] == str(gold["csv_row3_text"]) and '"quoted" tail' in csv[3]["text"]
    assert csv[np.float64(2.0)]["id"] == "3"
    bad = CSVDataset(os.path.join(DATA, "malformed.tsv"))
    assert bool(gold["csv_malformed_is_none"]) and bad[1] is None and bad[0]["id"] == "1"
    q = QueryCSVDataset(os.path.join(DATA, "questions.csv"))
    assert isinstance(q[0]["answers"], list) and len(q[0]["answers"]) == 2
    for k in [k for k in gold.files if k.startswith("test_sampler/")]:
        _, world, rank = k.split("/")
        assert contiguous_test_shard(len(csv), int(world), int(rank)) == gold[k].tolist(), k


def test_rank_shards_of_passages_datamodule(model_dir):
    """With a multi-rank trainer every rank loads its contiguous, unpadded slice (dpr.py:463-470)."""
    import types

    from dpr_scale_b200.datamodule.dpr import DenseRetrieverPassagesDataModule
    tf = HFTransform(model_path=model_dir, max_seq_len=24)
    seen = []
    for rank in range(3):
        dm = DenseRetrieverPassagesDataModule(tf, os.path.join(DATA, "passages.tsv"), test_batch_size=2,
                                              device_prefetch=False)
        dm.trainer = types.SimpleNamespace(world_size=3, global_rank=rank)
        seen.append([i for b in dm.test_dataloader() for i in b["corpus_ids"]])
    assert seen == [["1", "2", "3", "4"], ["5", "6", "7", "8"], ["9", "10", "11"]]


@pytest.mark.parametrize("prefetch", [0, 2])
def test_multi_jsonl_corpus_datamodule_equals_reference(gold, model_dir, prefetch):
    """DRAGON-style input (dragon_aws.yaml): several training files drawn per row with random.choice, contexts fetched
    from a corpus table by docidx (datamodule/dpr.py:333-412)."""
    import random

    from dpr_scale_b200.datamodule.dpr import DenseRetrieverMultiJsonlDataModule
    tf = HFTransform(model_path=model_dir, max_seq_len=24)
    la, lb = os.path.join(DATA, "light_a.jsonl"), os.path.join(DATA, "light_b.jsonl")
    dm = DenseRetrieverMultiJsonlDataModule(transform=tf, train_path=[la, lb], val_path=la, test_path=lb,
                                            corpus_path=os.path.join(DATA, "corpus.tsv"), batch_size=3, num_negative=2,
                                            pos_ctx_sample=True, num_val_negative=1, num_test_negative=3, use_title=True,
                                            prefetch_batches=prefetch, device_prefetch=False)
    assert len(dm.datasets["train"]) == 7
    random.seed(5)
    np.random.seed(99)
    for stage, loader in (("train", dm.train_dataloader()), ("valid", dm.val_dataloader()), ("test", dm.test_dataloader())):
        n = 0
        for i, batch in enumerate(loader):
            _same(batch, gold, f"multi/{stage}/{i}")
            n += 1
        assert n == int(gold[f"multi/{stage}/num_batches"])


def test_stream_worker_exits_promptly_when_abandoned_with_full_queue():
    """Consumer stops after one batch while the producer has finished everything and sits on a full queue: closing
    the iterator must not wait for the join timeout."""
    import time
    s = BatchStream(list(range(6)), lambda: list(range(6)), 1, lambda r: {"x": torch.tensor(r)}, prefetch_batches=2)
    it = iter(s)
    next(it)
    time.sleep(0.5)                       # producer: queue full, remaining batches / "end" pending
    t0 = time.perf_counter()
    it.close()
    assert time.perf_counter() - t0 < 2.0


def test_line_index_matches_readlines_on_random_files(tmp_path):
    """LineFile == file.readlines() (what the reference's readline loop yields) on random byte soups: empty lines,
    CRLF, unicode, with and without a trailing newline, spanning several scan windows."""
    import dpr_scale_b200.datamodule.dpr as D
    rng = np.random.RandomState(3)
    alphabet = ["a", "b", " ", "\t", "\r", "é", "漢", '"', "\n", "\n"]
    old = D._SCAN_BYTES
    D._SCAN_BYTES = 37                                     # force many scan windows
    try:
        for trial in range(40):
            text = "".join(rng.choice(alphabet, size=rng.randint(0, 400)))
            p = tmp_path / f"f{trial}.txt"
            p.write_bytes(text.encode())
            want = open(p, "rb").readlines()
            ds = LineFile(str(p))
            assert len(ds) == len(want), (trial, text)
            assert [ds[i] for i in range(len(ds))] == want, trial
            if want:
                hdr = LineFile(str(p), header=True)
                assert [hdr[i] for i in range(len(hdr))] == want[1:], trial
    finally:
        D._SCAN_BYTES = old
