"""CPU tests of the host-side plugin surface: config composition, `_target_` instantiation, C-ABI exports."""
import ctypes
import re
import os

from dpr_scale_b200 import _lib
from dpr_scale_b200.utils.config import compose, instantiate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_compose_defaults_and_overrides():
    cfg = compose("msmarco_baseline", ["task.model.model_path=/tmp/m", "task.warmup_steps=7", "+task.k=3"])
    assert cfg.task._target_ == "dpr_scale_b200.task.dpr_task.DenseRetrieverTask"
    assert cfg.task.model._target_ == "dpr_scale_b200.models.hf_model.HFEncoder"
    assert cfg.task.model.model_path == "/tmp/m"
    assert cfg.task.transform.text_transform.model_path == "/tmp/m"  # ${task.model.model_path} interpolation
    assert cfg.task.warmup_steps == 7 and cfg.task.k == 3
    assert cfg.task.optim.lr == 2.0e-05 and cfg.task.optim._target_.endswith("FusedAdamW")
    assert cfg.trainer.gradient_clip_val == 2.0 and cfg.datamodule.num_negative == 7


def test_instantiate_task_by_target_without_recursion():
    cfg = compose("config", [])
    cfg.task.datamodule = None
    task = instantiate(cfg.task, _recursive_=False)
    assert type(task).__name__ == "DenseRetrieverTask"
    assert task.model_conf["_target_"].endswith("HFEncoder")  # stays a config node until setup()
    assert task.shared_model is False and task.in_batch_negatives is True


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dprb.h")).read()
    declared = set(re.findall(r"\b(dprb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().dprb_version() == 100
