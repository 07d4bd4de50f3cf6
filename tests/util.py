"""Shared helpers for the parity tests (golden loading, synthetic models/batches)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def tokens_from(d, prefix):
    t = sub(d, prefix)
    return {k: v for k, v in t.items()}


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


BERT_TINY_CFG = {"layers": 2, "heads": 2, "ln_eps": 1e-12, "pad_id": 0, "roberta": False}
ROBERTA_TINY_CFG = {"layers": 2, "heads": 2, "ln_eps": 1e-5, "pad_id": 1, "roberta": True}
