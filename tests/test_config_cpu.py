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


def test_encoder_rejects_unsupported_configs_and_cpu_execution():
    import pytest
    import torch
    from dpr_scale_b200._lib import DprbError
    from dpr_scale_b200.models.hf_model import HFEncoder
    base = dict(vocab_size=64, hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256,
                max_position_embeddings=40)
    with pytest.raises(ValueError):   # head_dim must be 64
        HFEncoder.from_config({**base, "num_attention_heads": 4})
    with pytest.raises(ValueError):   # erf-GELU only
        HFEncoder.from_config({**base, "hidden_act": "relu"})
    with pytest.raises(ValueError):   # absolute position embeddings only
        HFEncoder.from_config({**base, "position_embedding_type": "relative_key"})
    enc = HFEncoder.from_config(base, dropout=0.0)
    with pytest.raises(DprbError):    # no CPU fallback: the product path fails loudly without a CUDA device
        enc({"input_ids": torch.ones(2, 4, dtype=torch.long)})
    with pytest.raises(FileNotFoundError):  # hub names cannot be resolved offline: must be a local directory
        HFEncoder(model_path="bert-base-uncased")


def test_state_dict_roundtrip_keeps_arena_views():
    import torch
    from dpr_scale_b200.models.hf_model import HFEncoder
    base = dict(vocab_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                max_position_embeddings=40)
    a, b = HFEncoder.from_config(base, seed=1), HFEncoder.from_config(base, seed=2)
    b.load_state_dict(a.state_dict())
    assert torch.equal(a.master, b.master)
    # parameters are views of ONE flat arena (Q, K, V adjacent => fused [3H, H] operand)
    lay = b.transformer.layout
    q_off = lay.by_name["encoder.layer.0.attention.self.query.weight"][1]
    k_off = lay.by_name["encoder.layer.0.attention.self.key.weight"][1]
    assert k_off - q_off == 128 * 128
    w = b.transformer.encoder.layer._modules["0"].attention.self.query.weight
    assert w.data_ptr() == b.master.data_ptr() + 4 * q_off
