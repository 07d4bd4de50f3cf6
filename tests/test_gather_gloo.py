"""world_size-2 gloo test (CPU) of the N>1 host logic: the packed all-gather of (q, c, labels, mask) must
reproduce the reference's gather semantics (dpr_task.py:163-195: rank-major concatenation, label offsets by the
number of contexts of the preceding ranks) — checked against the oracle and the 2-rank golden loss."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dpr_scale_b200.task.dpr_task import DenseRetrieverTask
    from oracle import encoder as oenc
    from oracle import task as otask
    from tests.util import BERT_TINY_CFG, load_golden, sub

    g2, g1 = load_golden("golden_2rank.npz"), load_golden("golden_1rank.npz")
    sd_q, sd_c = sub(g1, "sd_q/"), sub(g1, "sd_c/")
    T = float(g1["temperature"])
    reps = []
    for r in range(world):
        b = sub(g2, f"rank{r}/batch/")
        reps.append((oenc.encode(sd_q, BERT_TINY_CFG, sub(b, "query_ids/")),
                     oenc.encode(sd_c, BERT_TINY_CFG, sub(b, "contexts_ids/")), b["pos_ctx_indices"], b["ctx_mask"].bool()))
    q, c, lab, m = reps[rank]
    task = DenseRetrieverTask(transform={}, model={}, datamodule=None, optim={})
    q_all, c_all, labels, col_mask, q0, c0 = task._gather_global(q, c, lab, m)
    wq, wc, wl, wm = otask.gather_for_rank(rank, [x[0] for x in reps], [x[1] for x in reps], [x[2] for x in reps],
                                           [x[3] for x in reps])
    ok = (torch.equal(q_all, wq) and torch.equal(c_all, wc) and torch.equal(labels, wl)
          and torch.equal(col_mask.bool(), wm) and q0 == rank * q.shape[0] and c0 == rank * c.shape[0])
    loss, _ = otask.in_batch_loss(q_all, c_all, col_mask.bool(), labels, T)
    ret[rank] = (bool(ok), float(loss), float(g2[f"rank{rank}/loss"]))
    dist.barrier()
    dist.destroy_process_group()


def test_packed_gather_matches_reference_semantics_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, ret), nprocs=2, join=True)
    for r in (0, 1):
        ok, loss, want = ret[r]
        assert ok, f"rank {r}: gathered tensors differ from the oracle"
        assert abs(loss - want) < 1e-5, (loss, want)
