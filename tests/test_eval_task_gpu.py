"""generate_embeddings path (BASELINE config 5 shape, tiny): forward-only encoder mode + GenerateEmbeddingsTask writes a
pickle byte-compatible with the reference's `reps_{rank:04}.pkl` (protocol 4 of one fp32 tensor) whose values match the
oracle; also the validation metrics path of DenseRetrieverTask against the oracle's rank metrics."""
import pickle

import pytest
import torch

from tests.util import BERT_TINY_CFG, load_golden, rel_l2, sub

pytestmark = pytest.mark.gpu


def test_generate_embeddings_matches_oracle(tmp_path):
    from dpr_scale_b200.task.dpr_eval_task import GenerateEmbeddingsTask
    from oracle import encoder as oenc
    from tests.test_task_gpu import CFG, _batch
    g = load_golden("golden_1rank.npz")
    task = GenerateEmbeddingsTask(ctx_embeddings_dir=str(tmp_path), checkpoint_path="", transform={}, datamodule=None,
                                  optim={}, shared_model=False,
                                  model={"_target_": "dpr_scale_b200.models.hf_model.HFEncoder.from_config",
                                         "config": CFG, "dropout": 0.0})
    task.trainer = None
    task.setup("test")
    task.context_encoder.load_state_dict(sub(g, "sd_c/"))
    task = task.cuda().eval()
    batch = _batch(g)
    nb = 11                                           # more batches than pinned ring slots: slots are reused
    outs = [task.test_step({"contexts_ids": batch["contexts_ids"]}, i) for i in range(nb)]
    assert outs == [8] * nb
    path = task.test_epoch_end(outs)
    with open(path, "rb") as f:
        reps = pickle.load(f)
    assert reps.dtype == torch.float32 and tuple(reps.shape) == (8 * nb, 128) and not reps.is_cuda
    want = oenc.encode(sub(g, "sd_c/"), BERT_TINY_CFG, batch["contexts_ids"])
    for i in range(nb):
        assert rel_l2(reps[8 * i:8 * i + 8], want) <= 1e-2
    assert path.endswith("reps_0000.pkl")


def test_eval_metrics_match_oracle():
    from oracle import task as otask
    from tests.test_task_gpu import _batch, _task
    g = load_golden("golden_1rank.npz")
    task = _task(g).eval()
    task.in_batch_eval = True
    batch = _batch(g)
    with torch.no_grad():
        out = task.validation_step(batch, 0)
        metrics = task.validation_epoch_end([out])
    m = batch["ctx_mask"].repeat(4, 1)
    scores = otask.sim_score(g["q_emb"], g["c_emb"], m)
    rank, mrr, hit = otask.rank_metrics(scores, batch["pos_ctx_indices"], k=1)
    assert abs(float(metrics["valid_avg_rank"]) - rank / 4) < 1e-6
    assert abs(float(metrics["valid_mrr"]) - mrr / 4) < 1e-6
    assert abs(float(metrics["valid_accuracy@1"]) - hit / 4) < 1e-6
    want_loss = torch.nn.functional.cross_entropy(scores, batch["pos_ctx_indices"])
    assert abs(float(metrics["valid_loss"]) - float(want_loss)) < 5e-2
