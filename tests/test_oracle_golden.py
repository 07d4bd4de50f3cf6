"""Pins the oracle (oracle/*.py CPU restatement) against golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py ran /root/reference's HFEncoder + DenseRetrieverTask)."""
import torch

from oracle import encoder as oenc
from oracle import task as otask
from tests.util import BERT_TINY_CFG, ROBERTA_TINY_CFG, load_golden, sub


def _req(sd):
    return {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}


def test_encoder_and_loss_match_reference_1rank():
    g = load_golden("golden_1rank.npz")
    sd_q, sd_c = _req(sub(g, "sd_q/")), _req(sub(g, "sd_c/"))
    q = oenc.encode(sd_q, BERT_TINY_CFG, sub(g, "batch/query_ids/"))
    c = oenc.encode(sd_c, BERT_TINY_CFG, sub(g, "batch/contexts_ids/"))
    assert torch.allclose(q, g["q_emb"], atol=2e-5, rtol=1e-5)
    assert torch.allclose(c, g["c_emb"], atol=2e-5, rtol=1e-5)
    loss, logits = otask.in_batch_loss(q, c, g["batch/ctx_mask"], g["batch/pos_ctx_indices"], float(g["temperature"]))
    assert torch.allclose(logits, g["logits"], atol=1e-3, rtol=1e-5)
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    loss.backward()
    for name, sd in (("q", sd_q), ("c", sd_c)):
        grads = sub(g, f"grad_{name}/")
        assert len(grads) > 30
        for k, ref in grads.items():
            got = sd[k].grad
            assert got is not None, k
            assert torch.allclose(got, ref, atol=2e-5 + 1e-4 * float(ref.abs().max()), rtol=1e-4), k
    # the HF pooler receives no gradient in the reference either
    assert not any("pooler" in k for k in sub(g, "grad_q/"))


def test_gather_semantics_match_reference_2rank():
    g = load_golden("golden_2rank.npz")
    g1 = load_golden("golden_1rank.npz")
    sd_q0, sd_c0 = sub(g1, "sd_q/"), sub(g1, "sd_c/")
    for rank in (0, 1):
        sd_q, sd_c = _req(sd_q0), _req(sd_c0)
        qs, cs, labs, masks = [], [], [], []
        for r in (0, 1):
            b = sub(g, f"rank{r}/batch/")
            qs.append(oenc.encode(sd_q, BERT_TINY_CFG, sub(b, "query_ids/")))
            cs.append(oenc.encode(sd_c, BERT_TINY_CFG, sub(b, "contexts_ids/")))
            labs.append(b["pos_ctx_indices"])
            masks.append(b["ctx_mask"])
        q, c, lab, m = otask.gather_for_rank(rank, qs, cs, labs, masks)
        loss, _ = otask.in_batch_loss(q, c, m, lab, float(g1["temperature"]))
        assert abs(float(loss) - float(g[f"rank{rank}/loss"])) < 1e-5
        loss.backward()
        for name, sd in (("q", sd_q), ("c", sd_c)):
            for k, ref in sub(g, f"rank{rank}/grad_{name}/").items():
                assert torch.allclose(sd[k].grad, ref, atol=2e-5 + 1e-4 * float(ref.abs().max()), rtol=1e-4), k


def test_roberta_position_ids_match_reference():
    g = load_golden("golden_roberta.npz")
    rep = oenc.encode(sub(g, "sd/"), ROBERTA_TINY_CFG, sub(g, "tokens/"))
    assert torch.allclose(rep, g["rep"], atol=2e-5, rtol=1e-5)


def test_lr_lambda_and_adamw_match_torch():
    assert otask.lr_lambda(0, 10, 100) == 0.0
    assert otask.lr_lambda(5, 10, 100) == 0.5
    assert abs(otask.lr_lambda(55, 10, 100) - 0.5) < 1e-12
    assert otask.lr_lambda(200, 10, 100) == 0.0
    torch.manual_seed(0)
    p = torch.randn(1000)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.AdamW([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(1000)
        ref.grad = g.clone()
        opt.step()
        otask.adamw_step(p, g, m, v, step, 1e-3, weight_decay=0.01)
    assert torch.allclose(p, ref.detach(), atol=1e-6)


def test_rank_metrics():
    s = torch.tensor([[0.1, 0.9, 0.5], [0.8, 0.1, 0.3]])
    rank, mrr, hit = otask.rank_metrics(s, torch.tensor([2, 0]), k=1)
    assert (rank, hit) == (3, 1) and abs(mrr - 1.5) < 1e-9
