"""BertEncoder — mirror of ``dpr_scale.models.hf_bert.BertEncoder`` (/root/reference/dpr_scale/models/hf_bert.py:12-28):
BERT + CLS pooling.  (The reference class only runs on transformers 3.x — it tuple-unpacks the model output,
hf_bert.py:26; the arithmetic is identical to HFEncoder without a projection head.)"""
from .hf_model import HFEncoder


class BertEncoder(HFEncoder):
    def __init__(self, model_path: str = "bert-base-uncased", dropout: float = 0.1):
        super().__init__(model_path=model_path, dropout=dropout, projection_dim=None)
