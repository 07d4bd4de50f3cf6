"""world_size-2 gloo test (CPU) of the multi-GPU retrieval host logic (dpr_scale_b200/run_retrieval.py): file
ownership, global row offsets and the rank-major gather layout that feeds the merge - the sequential-shard layout
of run_retrieval_pytorch.py:218-227 with ranks in place of shards.  The GPU kernels are not involved."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dpr_scale_b200 import run_retrieval as RR
    from oracle import retrieval as R
    g = np.load(os.path.join(ROOT, "tests", "golden", "retrieval_small.npz"))
    files = [f"reps_{i:04}.pkl" for i in range(4)]
    mine = RR.rank_files(files, rank, world)
    shards = [g["shard0"], g["shard1"][:650]]              # unequal shard sizes
    off = RR.global_row_offset(len(shards[rank]), "cpu")
    k = int(g["topk"])
    s, i = R.search_index(g["queries"], shards[rank], 8, k, fp16_scores=False)
    gs, gi = RR.gather_rank_lists(torch.tensor(s, dtype=torch.float32), torch.tensor(i + off))
    ms, order = R.topk_desc(gs.numpy().astype(np.float64), k)
    mi = np.take_along_axis(gi.numpy(), order, axis=1)
    ws, wi = R.search_index(g["queries"], np.concatenate(shards, 0), 8, k, fp16_scores=False)
    ret[rank] = (mine, off, bool(np.allclose(ms, ws, rtol=0, atol=0)), bool(np.array_equal(mi, wi)), tuple(gs.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_sharded_retrieval_layout_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29633, ret), nprocs=2, join=True)
    assert ret[0][0] == ["reps_0000.pkl", "reps_0001.pkl"] and ret[1][0] == ["reps_0002.pkl", "reps_0003.pkl"]
    assert ret[0][1] == 0 and ret[1][1] == 700
    for r in (0, 1):
        assert ret[r][2] and ret[r][3], "merged rank lists differ from the single-index search"
        assert ret[r][4] == (23, 20)
