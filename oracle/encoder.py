"""fp32 CPU restatement of the encoder arithmetic the reference delegates to HuggingFace.

Follows, function by function:
  embeddings       site-packages/transformers/models/bert/modeling_bert.py:72-112 (BertEmbeddings.forward)
  roberta_pos_ids  site-packages/transformers/models/roberta/modeling_roberta.py:146-159
  self_attention   modeling_bert.py:168-207 (BertSelfAttention.forward; SDPA == softmax(QK^T/sqrt(dh)+mask)V)
  layer            modeling_bert.py:287-298, 330-356, 359-421 (BertSelfOutput, BertIntermediate, BertOutput, BertLayer)
  encode           modeling_bert.py:628-691 (BertModel.forward) + /root/reference/dpr_scale/models/hf_model.py:36-41
Written with plain torch ops on CPU tensors so autograd provides the reference gradients.
Parameter names are the HF state_dict names (prefix ``transformer.`` as in HFEncoder).
"""
import math

import torch


def roberta_position_ids(input_ids, pad_id):
    mask = (input_ids != pad_id).to(torch.int64)
    return torch.cumsum(mask, dim=1) * mask + pad_id


def layer_norm(x, g, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def gelu_erf(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def embeddings(sd, pfx, input_ids, token_type_ids, position_ids, eps, drop=None):
    e = sd[pfx + "embeddings.word_embeddings.weight"][input_ids]
    e = e + sd[pfx + "embeddings.token_type_embeddings.weight"][token_type_ids]
    e = e + sd[pfx + "embeddings.position_embeddings.weight"][position_ids]
    y = layer_norm(e, sd[pfx + "embeddings.LayerNorm.weight"], sd[pfx + "embeddings.LayerNorm.bias"], eps)
    return y if drop is None else y * drop  # drop = keep / (1 - p): modeling_bert.py:111


def self_attention(x, sd, lp, heads, attention_mask, drop=None):
    N, S, H = x.shape
    dh = H // heads
    a = lp + "attention.self."

    def proj(name):
        y = x @ sd[a + name + ".weight"].T + sd[a + name + ".bias"]
        return y.view(N, S, heads, dh).transpose(1, 2)  # N, A, S, dh

    q, k, v = proj("query"), proj("key"), proj("value")
    scores = q @ k.transpose(-1, -2) / math.sqrt(dh)
    if attention_mask is not None:
        neg = torch.zeros(N, 1, 1, S, dtype=x.dtype)
        neg = neg.masked_fill(attention_mask.view(N, 1, 1, S) == 0, float("-inf"))
        scores = scores + neg
    p = torch.softmax(scores, dim=-1)
    if drop is not None:  # attention-probability dropout (eager_attention_forward / SDPA dropout_p)
        p = p * drop
    ctx = (p @ v).transpose(1, 2).reshape(N, S, H)
    return ctx


def layer(x, sd, lp, heads, attention_mask, eps, drops=None):
    """drops: optional dict with multiplier tensors (keep / (1-p)) under 'attn', 'attn_out', 'ffn_out'."""
    drops = drops or {}
    ctx = self_attention(x, sd, lp, heads, attention_mask, drops.get("attn"))
    o = ctx @ sd[lp + "attention.output.dense.weight"].T + sd[lp + "attention.output.dense.bias"]
    if "attn_out" in drops:
        o = o * drops["attn_out"]  # BertSelfOutput.dropout, modeling_bert.py:296
    x1 = layer_norm(o + x, sd[lp + "attention.output.LayerNorm.weight"], sd[lp + "attention.output.LayerNorm.bias"], eps)
    h = gelu_erf(x1 @ sd[lp + "intermediate.dense.weight"].T + sd[lp + "intermediate.dense.bias"])
    o2 = h @ sd[lp + "output.dense.weight"].T + sd[lp + "output.dense.bias"]
    if "ffn_out" in drops:
        o2 = o2 * drops["ffn_out"]  # BertOutput.dropout, modeling_bert.py:354
    return layer_norm(o2 + x1, sd[lp + "output.LayerNorm.weight"], sd[lp + "output.LayerNorm.bias"], eps)


def encode(sd, cfg, tokens, prefix="transformer.", dropout=None):
    """tokens: mapping with input_ids (+ optional token_type_ids, attention_mask) -> CLS reps [N, H].

    cfg keys: layers, heads, ln_eps, pad_id, roberta (bool).
    """
    input_ids = tokens["input_ids"]
    N, S = input_ids.shape
    tt = tokens.get("token_type_ids")
    if tt is None:
        tt = torch.zeros_like(input_ids)
    am = tokens.get("attention_mask")
    if cfg.get("roberta", False):
        pos = roberta_position_ids(input_ids, cfg["pad_id"])
    else:
        pos = torch.arange(S).unsqueeze(0).expand(N, S)
    dropout = dropout or {}
    x = embeddings(sd, prefix, input_ids, tt, pos, cfg["ln_eps"], dropout.get("emb"))
    for l in range(cfg["layers"]):
        x = layer(x, sd, f"{prefix}encoder.layer.{l}.", cfg["heads"], am, cfg["ln_eps"], dropout.get(l))
    rep = x[:, 0, :]
    if prefix.replace("transformer.", "project.0.weight") in sd:  # optional Linear+LayerNorm projection head
        pp = prefix.replace("transformer.", "project.")
        rep = rep @ sd[pp + "0.weight"].T + sd[pp + "0.bias"]
        rep = layer_norm(rep, sd[pp + "1.weight"], sd[pp + "1.bias"], 1e-5)
    return rep.clone()
