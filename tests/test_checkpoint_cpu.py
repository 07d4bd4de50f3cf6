"""ModelCheckpoint stand-in (dpr_scale_b200/utils/checkpoint.py): top-k by a monitored metric, save_last, Lightning-style
file versioning, atomic writes, reloadable state_dict with the reference's key layout."""
import os

import torch

from dpr_scale_b200.utils.checkpoint import ModelCheckpoint, load_into


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.query_encoder = torch.nn.Linear(3, 2)
        self.hparams = {"k": 1}
        self.loaded = 0

    def on_load_checkpoint(self, ckpt):
        self.loaded += 1


def test_top_k_by_metric_and_last(tmp_path):
    cb = ModelCheckpoint(dirpath=str(tmp_path), monitor="valid_mrr", mode="max", save_last=True, save_top_k=2,
                         filename="checkpoint_best")
    m = Tiny()
    scores = [0.2, 0.5, 0.1, 0.7, 0.6]
    for epoch, s in enumerate(scores):
        with torch.no_grad():
            m.query_encoder.weight.fill_(s)
        cb.on_validation_end(m, epoch, 10 * (epoch + 1), {"valid_mrr": torch.tensor(s), "valid_loss": 1.0})
    kept = sorted(f for f in os.listdir(tmp_path) if f != "last.ckpt")
    assert len(kept) == 2 and all(f.startswith("checkpoint_best") for f in kept)
    assert abs(cb.best_model_score - 0.7) < 1e-6 and [round(v, 5) for v in sorted(cb.best_k.values())] == [0.6, 0.7]
    best = torch.load(cb.best_model_path, weights_only=False)
    assert abs(float(best["state_dict"]["query_encoder.weight"][0, 0]) - 0.7) < 1e-6 and best["epoch"] == 3 and best["global_step"] == 40
    last = torch.load(os.path.join(tmp_path, "last.ckpt"), weights_only=False)
    assert abs(float(last["state_dict"]["query_encoder.weight"][0, 0]) - 0.6) < 1e-6      # last epoch, even though not the best
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".tmp")]
    fresh = Tiny()
    ck = load_into(fresh, cb.best_model_path)
    assert fresh.loaded == 1 and abs(float(fresh.query_encoder.weight[0, 0]) - 0.7) < 1e-6 and ck["hyper_parameters"] == {"k": 1}


def test_min_mode_missing_metric_and_non_writer_rank(tmp_path):
    cb = ModelCheckpoint(dirpath=str(tmp_path / "a"), monitor="valid_loss", mode="min", save_top_k=1)
    m = Tiny()
    assert cb.on_validation_end(m, 0, 1, {"other": 1.0}) == []             # monitored metric absent: nothing saved
    cb.on_validation_end(m, 0, 1, {"valid_loss": 2.0})
    first = cb.best_model_path
    cb.on_validation_end(m, 1, 2, {"valid_loss": 3.0})                      # worse: not saved
    assert cb.best_model_path == first and len(os.listdir(tmp_path / "a")) == 1
    cb.on_validation_end(m, 2, 3, {"valid_loss": 1.0})
    assert cb.best_model_path != first and not os.path.exists(first) and cb.best_model_score == 1.0
    assert os.path.basename(cb.best_model_path) == "epoch=2-step=3.ckpt"
    other = ModelCheckpoint(dirpath=str(tmp_path / "b"), monitor=None, save_top_k=1, save_last=True)
    other.on_validation_end(m, 0, 5, None, is_writer=False)                 # ranks != 0 track paths but write nothing
    assert other.best_model_path.endswith("epoch=0-step=5.ckpt") and not os.path.exists(tmp_path / "b")


class StubTask(torch.nn.Module):
    """Lightning-hook-shaped stand-in (no encoders): lets the trainer's loop logic run on the CPU."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor([4.0]))
        self.trainer = None
        self.setup_done = False
        self.val_calls = 0

    def setup(self, stage):
        self.setup_done = True

    def configure_optimizers(self):
        opt = torch.optim.SGD(self.parameters(), lr=0.25)
        return [opt], [{"scheduler": torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0), "interval": "step"}]

    def training_step(self, batch, idx):
        return (self.w * batch["x"]).pow(2).sum()

    def validation_step(self, batch, idx):
        return float(batch["x"].sum())

    def validation_epoch_end(self, outs):
        self.val_calls += 1
        return {"valid_mrr": torch.tensor(1.0 / (1.0 + abs(float(self.w)))), "n": len(outs)}

    def test_step(self, batch, idx):
        return float(self.w)

    def test_epoch_end(self, outs):
        return {"w_seen": outs[0], "n": len(outs)}


class StubData:
    trainer = None

    def __init__(self):
        self.epochs = []

    def set_epoch(self, e):
        self.epochs.append(e)

    def train_dataloader(self):
        return [{"x": torch.tensor([1.0])} for _ in range(3)]

    def val_dataloader(self):
        return [{"x": torch.tensor([1.0])} for _ in range(2)]

    test_dataloader = val_dataloader


def test_fit_validates_checkpoints_and_tests_best(tmp_path):
    from dpr_scale_b200.trainer import Trainer
    cb = ModelCheckpoint(dirpath=str(tmp_path), monitor="valid_mrr", mode="max", save_last=True, save_top_k=1,
                         filename="checkpoint_best")
    task, data = StubTask(), StubData()
    tr = Trainer(max_epochs=3, device="cpu", callbacks=[cb], log_every_n_steps=1000)
    tr.fit(task, data)
    assert data.epochs == [0, 1, 2] and task.val_calls == 3 and tr.global_step == 9 and data.trainer is tr
    assert abs(float(task.w)) < 4.0                              # SGD on w^2 shrinks |w| every step
    files = sorted(os.listdir(tmp_path))                           # best = last epoch (smallest |w|); older ones pruned
    assert len(files) == 2 and files[0].startswith("checkpoint_best") and files[1] == "last.ckpt"
    best_w = float(torch.load(cb.best_model_path, weights_only=False)["state_dict"]["w"])
    assert abs(best_w - float(task.w)) < 1e-7
    with torch.no_grad():
        task.w.fill_(123.0)
    res = tr.test(None, data, ckpt_path="best")                  # main.py:46-47: test the best checkpoint
    assert abs(res["w_seen"] - best_w) < 1e-7 and res["n"] == 2
    tr2 = Trainer(max_epochs=5, max_steps=4, device="cpu", check_val_every_n_epoch=0)
    t2 = StubTask()
    tr2.fit(t2, StubData())
    assert tr2.global_step == 4 and t2.val_calls == 0
