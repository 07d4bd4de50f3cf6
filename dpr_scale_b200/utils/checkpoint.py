"""Checkpoint callback for the mini trainer with the knobs of ``pytorch_lightning.callbacks.ModelCheckpoint`` that the
reference configures (/root/reference/dpr_scale/conf/checkpoint_callback/default.yaml: monitor valid_mrr, mode max,
save_last, save_top_k 3, filename checkpoint_best) and main.py:30-31,46-47 uses (``best_model_path``).

Files are Lightning-shaped dicts - ``{"state_dict", "epoch", "global_step", ...}`` with the task's parameter names
(``query_encoder.transformer.*`` / ``context_encoder.transformer.*``) - so they load into the reference's
``DenseRetrieverTask`` and into ``GenerateEmbeddingsTask.setup`` (dpr_eval_task.py:26-30) alike.
"""
import os

import torch
import torch.distributed as dist


class ModelCheckpoint:
    def __init__(self, dirpath=None, monitor=None, mode="min", save_last=None, save_top_k=1, filename=None,
                 verbose=False, **unused):
        assert mode in ("min", "max"), mode
        self.dirpath = dirpath or os.path.join(os.getcwd(), "checkpoints")
        self.monitor, self.mode, self.save_last = monitor, mode, bool(save_last)
        self.save_top_k, self.filename, self.verbose = int(save_top_k), filename, verbose
        self.best_k = {}                 # path -> score
        self.best_model_path, self.best_model_score, self.last_model_path = "", None, ""

    # ------------------------------------------------------------------ helpers
    def _better(self, a, b):
        return a > b if self.mode == "max" else a < b

    def _new_path(self, epoch, step):
        stem = self.filename or f"epoch={epoch}-step={step}"
        stem = stem.format(epoch=epoch, step=step)
        path, v = os.path.join(self.dirpath, stem + ".ckpt"), 0
        while os.path.exists(path):      # Lightning's "-v1", "-v2" versioning of a fixed file name
            v += 1
            path = os.path.join(self.dirpath, f"{stem}-v{v}.ckpt")
        return path

    def _agreed_path(self, epoch, step):
        """One process per GPU: only rank 0 probes the directory (the others would race with its write and could pick
        a `-vN` name that never exists); the choice is broadcast so `best_k` / `best_model_path` agree on all ranks."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return self._new_path(epoch, step)
        box = [self._new_path(epoch, step) if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    @staticmethod
    def _payload(task, epoch, step, extra=None):
        sd = {k: v.detach().to("cpu", copy=True) for k, v in task.state_dict().items()}
        out = {"state_dict": sd, "epoch": int(epoch), "global_step": int(step), "pytorch-lightning_version": "1.6.4",
               "hyper_parameters": dict(getattr(task, "hparams", {}) or {})}
        out.update(extra or {})
        return out

    def _write(self, payload, path):
        os.makedirs(self.dirpath, exist_ok=True)
        tmp = path + ".tmp"
        torch.save(payload, tmp)
        os.replace(tmp, path)            # a killed job never leaves a truncated checkpoint behind
        if self.verbose:
            print(f"Saved checkpoint {path}")

    # ------------------------------------------------------------------ hook
    def on_validation_end(self, task, epoch, step, metrics, is_writer=True):
        """Called by the trainer after every validation pass with the logged metrics; returns the paths written."""
        written = []
        score = None
        if self.monitor is not None and metrics is not None and self.monitor in metrics:
            score = float(metrics[self.monitor])
        payload = None
        if self.save_top_k != 0 and (self.monitor is None or score is not None):
            worst = None
            if self.monitor is not None and self.save_top_k > 0 and len(self.best_k) >= self.save_top_k:
                worst = min(self.best_k, key=self.best_k.get) if self.mode == "max" else max(self.best_k, key=self.best_k.get)
                if not self._better(score, self.best_k[worst]):
                    worst = "skip"
            if worst != "skip":
                path = self._agreed_path(epoch, step)
                if is_writer:
                    payload = self._payload(task, epoch, step, {"monitor": self.monitor, "score": score})
                    self._write(payload, path)
                if self.monitor is None:
                    for old in list(self.best_k):            # no metric: keep only the newest save_top_k files
                        if self.save_top_k > 0 and len(self.best_k) >= self.save_top_k:
                            self.best_k.pop(old)
                            if is_writer and os.path.exists(old):
                                os.remove(old)
                    self.best_k[path] = float(step)
                    self.best_model_path, self.best_model_score = path, None
                else:
                    if worst is not None:
                        self.best_k.pop(worst)
                        if is_writer and os.path.exists(worst):
                            os.remove(worst)
                    self.best_k[path] = score
                    best = max(self.best_k, key=self.best_k.get) if self.mode == "max" else min(self.best_k, key=self.best_k.get)
                    self.best_model_path, self.best_model_score = best, self.best_k[best]
                written.append(path)
        if self.save_last:
            path = os.path.join(self.dirpath, "last.ckpt")
            if is_writer:
                self._write(payload or self._payload(task, epoch, step), path)
            self.last_model_path = path
            written.append(path)
        return written


def load_into(task, path):
    """Load a checkpoint written above (or by Lightning) into ``task``; returns the checkpoint dict."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(task, "on_load_checkpoint") and not getattr(task, "setup_done", False):
        task.on_load_checkpoint(ckpt)      # builds the encoders (dpr_task.py:78-79); skipped when they already exist
    task.load_state_dict(ckpt["state_dict"])
    return ckpt
