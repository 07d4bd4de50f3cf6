"""GenerateEmbeddingsTask — mirror of ``dpr_scale.task.dpr_eval_task.GenerateEmbeddingsTask``
(/root/reference/dpr_scale/task/dpr_eval_task.py:13-49): forward-only context encoding of a contiguous corpus shard,
written as ``reps_{rank:04}.pkl`` (pickle protocol 4 of one fp32 tensor, byte-compatible with the reference's output).

Differences that matter on a B200: the encoder runs in ``save_for_backward=0`` mode (two activation slots instead
of L), results are copied to a pinned host buffer asynchronously (no per-batch ``.cpu()`` sync) and concatenated once.
"""
import os
import pathlib
import pickle

import torch
import torch.distributed as dist

from .dpr_task import DenseRetrieverTask


class GenerateEmbeddingsTask(DenseRetrieverTask):
    def __init__(self, ctx_embeddings_dir, checkpoint_path, **kwargs):
        super().__init__(**kwargs)
        self.ctx_embeddings_dir = ctx_embeddings_dir
        self.checkpoint_path = checkpoint_path
        pathlib.Path(ctx_embeddings_dir).mkdir(parents=True, exist_ok=True)

    def setup(self, stage: str):
        super().setup("train")
        if self.checkpoint_path:
            print(f"Loading checkpoint from {self.checkpoint_path}")
            ckpt = torch.load(self.checkpoint_path, map_location="cpu", weights_only=False)
            self.load_state_dict(ckpt["state_dict"])

    def forward(self, contexts_ids):
        return self.encode_contexts(contexts_ids)

    @staticmethod
    def _to_pinned(rep):
        """Asynchronous D2H into a pinned buffer (the reference syncs with ``.cpu()`` every batch, :35)."""
        host = torch.empty(rep.shape, dtype=rep.dtype, pin_memory=rep.is_cuda)
        host.copy_(rep, non_blocking=True)
        return host

    @staticmethod
    def _collect(parts):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return torch.cat(parts, dim=0)

    @torch.no_grad()
    def _eval_step(self, batch, batch_idx):
        return self._to_pinned(self(batch["contexts_ids"]))

    def test_step(self, batch, batch_idx):
        return self._eval_step(batch, batch_idx)

    def test_epoch_end(self, contexts_repr):
        contexts_repr = self._collect(contexts_repr)
        if not self.ctx_embeddings_dir:
            self.ctx_embeddings_dir = getattr(self.trainer, "weights_save_path", ".")
        out_file = os.path.join(self.ctx_embeddings_dir, f"reps_{self.global_rank:04}.pkl")
        print(f"\nWriting tensor of size {contexts_repr.size()} to {out_file}")
        with open(out_file, mode="wb") as f:
            pickle.dump(contexts_repr, f, protocol=4)
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
        return out_file


class GenerateQueryEmbeddingsTask(GenerateEmbeddingsTask):
    """Mirror of ``GenerateQueryEmbeddingsTask`` (/root/reference/dpr_scale/task/dpr_eval_task.py:52-84): encode the
    question file with the query encoder and write one fp32 tensor to ``query_emb_output_path`` (default
    ``<ctx_embeddings_dir>/query_reps.pkl``, the file run_retrieval reads)."""

    def __init__(self, hnsw_index=False, output_path="/tmp/results.jsonl", query_emb_output_path=None, passages="",
                 **kwargs):
        super().__init__(**kwargs)
        self.hnsw_index = hnsw_index
        self.output_path = output_path
        self.query_emb_output_path = query_emb_output_path or os.path.join(self.ctx_embeddings_dir, "query_reps.pkl")

    def forward(self, query_ids):
        return self.encode_queries(query_ids)

    @torch.no_grad()
    def _eval_step(self, batch, batch_idx):
        return self._to_pinned(self(batch["query_ids"]))

    def test_epoch_end(self, queries_repr):
        queries_repr = self._collect(queries_repr)
        out_file = self.query_emb_output_path
        pathlib.Path(out_file).parent.mkdir(parents=True, exist_ok=True)
        print(f"\nWriting tensor of size {queries_repr.size()} to {out_file}")
        with open(out_file, mode="wb") as f:
            pickle.dump(queries_repr, f, protocol=4)
        return out_file
