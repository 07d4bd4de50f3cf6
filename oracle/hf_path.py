"""The reference's own CPU path, re-created without /root/reference: HuggingFace ``BertModel`` /
``RobertaModel`` (the third-party code dpr_scale/models/hf_model.py:25 instantiates) + CLS pooling
(hf_model.py:38-41) + the task arithmetic of oracle/task.py.  Used (a) to pin oracle/encoder.py,
(b) as the timed CPU baseline of bench.py (``cpu_baseline`` and ``--impl reference``).
transformers is a dependency of the reference (requirements.txt:6, pinned 3.4.0; 5.5.0 installed here).
"""
import torch.nn as nn


def make_config(kind="bert", **kw):
    from transformers import BertConfig, RobertaConfig
    if kind == "roberta":
        return RobertaConfig(**kw)
    return BertConfig(**kw)


class CLSEncoder(nn.Module):
    """AutoModel + last_hidden_state[:, 0, :].clone()  — the arithmetic of HFEncoder.forward."""

    def __init__(self, config, dropout=0.0, model=None):
        super().__init__()
        if model is not None:          # an already built (e.g. seeded, tests/realdims.py) HF module
            self.transformer = model
            return
        from transformers import AutoModel
        config.attention_probs_dropout_prob = dropout
        config.hidden_dropout_prob = dropout
        self.transformer = AutoModel.from_config(config)

    def forward(self, tokens):
        out = self.transformer(**tokens)
        return out[0][:, 0, :].clone()


def cfg_dict(config):
    return {
        "layers": config.num_hidden_layers,
        "heads": config.num_attention_heads,
        "ln_eps": config.layer_norm_eps,
        "pad_id": config.pad_token_id,
        "roberta": config.model_type == "roberta",
    }
