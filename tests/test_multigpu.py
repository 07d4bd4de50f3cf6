"""2-rank NCCL parity of the global in-batch-negative path (configs[2]) against the 2-rank golden produced by the
unmodified reference under gloo: every rank reports the reference's loss; after the trainer's gradient all-reduce
(SUM, scaled 1/W by the optimizer) the flat gradient equals the mean of the reference's per-rank gradients."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from dpr_scale_b200.task.dpr_task import DenseRetrieverTask
    from dpr_scale_b200.trainer import Trainer
    from tests.test_task_gpu import CFG, _batch
    from tests.util import load_golden, sub
    g2, g1 = load_golden("golden_2rank.npz"), load_golden("golden_1rank.npz")
    task = DenseRetrieverTask(transform={}, datamodule=None, shared_model=False, softmax_temperature=float(g1["temperature"]),
                              model={"_target_": "dpr_scale_b200.models.hf_model.HFEncoder.from_config", "config": CFG, "dropout": 0.0},
                              optim={"_target_": "dpr_scale_b200.optim.FusedAdamW", "lr": 0.0})
    tr = Trainer(max_steps=10, gradient_clip_val=0.0, device=torch.device("cuda", rank), grad_bucket_layers=1)
    tr.attach(task, None, "fit")
    task.query_encoder.load_state_dict(sub(g1, "sd_q/"))
    task.context_encoder.load_state_dict(sub(g1, "sd_c/"))
    task.train()
    batch = _batch(g2, f"rank{rank}/batch/")
    tr.optimizer.zero_grad()
    loss = task.training_step(batch, 0)
    loss.backward()
    tr._allreduce_grads()
    torch.cuda.synchronize()
    num = den = 0.0
    for name, enc in (("q", task.query_encoder), ("c", task.context_encoder)):
        r0, r1 = sub(g2, f"rank0/grad_{name}/"), sub(g2, f"rank1/grad_{name}/")
        for k, p in enc.named_parameters():
            if k not in r0:
                continue
            want = r0[k] + r1[k]  # SUM over ranks (the optimizer applies 1/W)
            got = p.grad.detach().float().cpu()
            num += float(((got - want).double() ** 2).sum())
            den += float((want.double() ** 2).sum())
    ret[rank] = (float(loss), float(g2[f"rank{rank}/loss"]), (num / den) ** 0.5, float(g1["amp_global_rel"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_global_negatives_match_reference():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29655, ret), nprocs=2, join=True)
    for r in (0, 1):
        loss, want, grel, amp = ret[r]
        assert abs(loss - want) <= 5e-2, (loss, want)
        assert grel <= 1.5 * amp, (grel, amp)
    assert abs(ret[0][0] - ret[1][0]) < 1e-6  # every rank computes the same global loss
