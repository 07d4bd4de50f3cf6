"""Input pipeline on the GPU: batches staged on the device by the background thread equal the reference's batches
(tests/golden/data_batches.npz), and the whole chain JSONL -> datamodule -> DenseRetrieverTask -> Trainer.fit runs,
with the first step's loss equal to the CPU oracle's loss for the reference's first batch (config 1 end to end)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "golden", "data")
JSONL = os.path.join(DATA, "synth.jsonl")
CASE_A = dict(batch_size=4, num_negative=3, neg_ctx_sample=True, pos_ctx_sample=False, num_val_negative=2,
              num_test_negative=5, use_title=True)


def _model_dir(tmp_path):
    from transformers import BertConfig
    vocab = open(os.path.join(DATA, "vocab.txt")).read()
    BertConfig(vocab_size=len(vocab.split()), hidden_size=16, num_hidden_layers=1, num_attention_heads=1,
               intermediate_size=16).save_pretrained(tmp_path)
    (tmp_path / "vocab.txt").write_text(vocab)
    return str(tmp_path)


def test_device_staged_batches_equal_reference(tmp_path):
    from dpr_scale_b200.datamodule.dpr import DenseRetrieverJsonlDataModule
    from dpr_scale_b200.transforms.hf_transform import HFTransform
    gold = np.load(os.path.join(HERE, "golden", "data_batches.npz"))
    tf = HFTransform(model_path=_model_dir(tmp_path), max_seq_len=24)
    dm = DenseRetrieverJsonlDataModule(transform=tf, train_path=JSONL, val_path=JSONL, test_path=JSONL,
                                       prefetch_batches=2, device_prefetch=True, **CASE_A)
    np.random.seed(1234)
    n = 0
    for i, batch in enumerate(dm.train_dataloader()):
        for key in ("query_ids", "contexts_ids"):
            for kk, v in batch[key].items():
                assert v.is_cuda and np.array_equal(v.cpu().numpy(), gold[f"a/train/{i}/{key}/{kk}"])
        for key in ("pos_ctx_indices", "scores", "ctx_mask"):
            assert batch[key].is_cuda and np.array_equal(batch[key].cpu().numpy(), gold[f"a/train/{i}/{key}"])
        n += 1
    assert n == int(gold["a/train/num_batches"])


def test_fit_from_jsonl_first_loss_matches_oracle(tmp_path):
    from dpr_scale_b200.datamodule.dpr import DenseRetrieverJsonlDataModule
    from dpr_scale_b200.task.dpr_task import DenseRetrieverTask
    from dpr_scale_b200.trainer import Trainer
    from dpr_scale_b200.transforms.hf_transform import HFTransform
    from oracle import encoder as oenc
    from oracle import task as otask
    from tests.test_task_gpu import CFG
    from tests.util import BERT_TINY_CFG
    gold = np.load(os.path.join(HERE, "golden", "data_batches.npz"))
    tf = HFTransform(model_path=_model_dir(tmp_path), max_seq_len=24)
    dm = DenseRetrieverJsonlDataModule(transform=tf, train_path=JSONL, val_path=JSONL, test_path=JSONL, **CASE_A)
    torch.manual_seed(0)
    task = DenseRetrieverTask(transform={}, datamodule=None, shared_model=False, softmax_temperature=8.0,
                              model={"_target_": "dpr_scale_b200.models.hf_model.HFEncoder.from_config", "config": CFG,
                                     "dropout": 0.0},
                              optim={"_target_": "dpr_scale_b200.optim.FusedAdamW", "lr": 1e-3}, warmup_steps=0)
    tr = Trainer(max_steps=4, gradient_clip_val=2.0, log_every_n_steps=1000)
    tr.attach(task, dm, "fit")
    sd_q = {k: v.detach().float().cpu().clone() for k, v in task.query_encoder.state_dict().items()}
    sd_c = {k: v.detach().float().cpu().clone() for k, v in task.context_encoder.state_dict().items()}
    task.train()
    np.random.seed(1234)
    losses = []
    for i, batch in enumerate(dm.train_dataloader()):
        losses.append(float(tr.training_step(batch, i)))
    assert len(losses) == int(gold["a/train/num_batches"]) and all(np.isfinite(losses))

    def tok(key):
        return {kk.split("/")[-1]: torch.from_numpy(gold[kk]) for kk in gold.files if kk.startswith(f"a/train/0/{key}/")}
    q = oenc.encode(sd_q, BERT_TINY_CFG, tok("query_ids"))
    c = oenc.encode(sd_c, BERT_TINY_CFG, tok("contexts_ids"))
    want, _ = otask.in_batch_loss(q, c, torch.from_numpy(gold["a/train/0/ctx_mask"]),
                                  torch.from_numpy(gold["a/train/0/pos_ctx_indices"]), 8.0)
    assert abs(losses[0] - float(want)) <= 5e-2, (losses[0], float(want))
