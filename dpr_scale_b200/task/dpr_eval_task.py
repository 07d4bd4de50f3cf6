"""Forward-only embedding dumps: ``GenerateEmbeddingsTask`` (passages -> ``reps_{rank:04}.pkl``) and
``GenerateQueryEmbeddingsTask`` (questions -> ``query_reps.pkl``), drop-ins for the classes of the same names in
/root/reference/dpr_scale/task/dpr_eval_task.py (:13-49 and :52-84): same constructor keywords, same Lightning test
hooks, same output files (pickle protocol 4 of ONE fp32 CPU tensor, which is what run_retrieval reads).

How a B200 changes the loop: the encoder runs in its forward-only mode (two activation slots instead of one per layer),
every batch result goes to a pinned host buffer with an asynchronous copy instead of the reference's blocking ``.cpu()``
per batch, and the device is synchronised once, right before the shard is concatenated and written.
"""
import os
import pathlib
import pickle

import torch
import torch.distributed as dist

from .dpr_task import DenseRetrieverTask


class _EmbeddingDumpTask(DenseRetrieverTask):
    """Shared machinery: which batch entry to encode, with which encoder, and where the shard goes."""

    batch_key = None

    def __init__(self, ctx_embeddings_dir, checkpoint_path, **task_kwargs):
        super().__init__(**task_kwargs)
        self.ctx_embeddings_dir = ctx_embeddings_dir
        self.checkpoint_path = checkpoint_path
        pathlib.Path(ctx_embeddings_dir).mkdir(parents=True, exist_ok=True)

    def setup(self, stage: str):
        super().setup("train")                       # always build the encoders, whatever stage the trainer names
        if not self.checkpoint_path:
            return
        print(f"Loading checkpoint from {self.checkpoint_path}")
        state = torch.load(self.checkpoint_path, map_location="cpu", weights_only=False)["state_dict"]
        self.load_state_dict(state)

    # -- per-batch: encode, then park the result in pinned memory without waiting for it
    def _encode(self, tokens):
        raise NotImplementedError

    def forward(self, tokens):
        return self._encode(tokens)

    @staticmethod
    def _to_pinned(rep):
        parked = torch.empty(rep.shape, dtype=rep.dtype, pin_memory=rep.is_cuda)
        parked.copy_(rep, non_blocking=True)
        return parked

    @torch.no_grad()
    def _eval_step(self, batch, batch_idx):
        return self._to_pinned(self(batch[self.batch_key]))

    def test_step(self, batch, batch_idx):
        return self._eval_step(batch, batch_idx)

    # -- per-shard: one sync, one concatenation, one pickle
    @staticmethod
    def _collect(parts):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return torch.cat(parts, dim=0)

    @staticmethod
    def _dump(tensor, out_file):
        pathlib.Path(out_file).parent.mkdir(parents=True, exist_ok=True)
        print(f"\nWriting tensor of size {tensor.size()} to {out_file}")
        with open(out_file, mode="wb") as f:
            pickle.dump(tensor, f, protocol=4)
        return out_file


class GenerateEmbeddingsTask(_EmbeddingDumpTask):
    """Passage side: context encoder over ``batch["contexts_ids"]``; every rank writes its own shard file."""

    batch_key = "contexts_ids"

    def _encode(self, contexts_ids):
        return self.encode_contexts(contexts_ids)

    def test_epoch_end(self, contexts_repr):
        shard = self._collect(contexts_repr)
        if not self.ctx_embeddings_dir:
            self.ctx_embeddings_dir = getattr(self.trainer, "weights_save_path", ".")
        out_file = self._dump(shard, os.path.join(self.ctx_embeddings_dir, f"reps_{self.global_rank:04}.pkl"))
        if dist.is_available() and dist.is_initialized():
            dist.barrier()                           # nobody leaves before every shard is on disk (:49)
        return out_file


class GenerateQueryEmbeddingsTask(GenerateEmbeddingsTask):
    """Question side: query encoder over ``batch["query_ids"]``; one file, by default next to the passage shards."""

    batch_key = "query_ids"

    def __init__(self, hnsw_index=False, output_path="/tmp/results.jsonl", query_emb_output_path=None, passages="",
                 **kwargs):
        super().__init__(**kwargs)
        self.hnsw_index = hnsw_index
        self.output_path = output_path
        self.query_emb_output_path = query_emb_output_path or os.path.join(self.ctx_embeddings_dir, "query_reps.pkl")

    def _encode(self, query_ids):
        return self.encode_queries(query_ids)

    def test_epoch_end(self, queries_repr):
        return self._dump(self._collect(queries_repr), self.query_emb_output_path)
