"""Parity of the ASSEMBLED CUDA path at BASELINE.json's real model dimensions against goldens the
unmodified reference produced (tests/golden/make_golden_realdims.py) and against the HF oracle run live on the host:

  * BERT-base H768 / L12 / A12 on config 1's shape (8 q + 16 ctx, S = 64, padded): embeddings, logits, loss, gradients
    of one full `training_step`, plus a well-conditioned linear-probe gradient check;
  * RoBERTa-large H1024 / L24 / A16, S = 256, pad-derived position ids (config 4's model): same;
  * the bench batch at full size (T = 131 072 context tokens): 16 of the 1024 sequences against the oracle, forward
    AND backward - catches 32-bit offset overflow / a wrong 12-layer assembly that property tests cannot see;
  * `shared_model=True` (the reference's constructor default) with query and context passes of EQUAL shape;
  * the `projection_dim` head.

Gates (vs the fp32 reference): embeddings rel-L2 <= 1e-2; logits max-abs <= 1e-2 * max|logit|; the loss
kernel within 2e-3 of the reference's cross-entropy formula evaluated on our own logits, and within max(5e-2, 2 x the
measured max|dlogit|) of the reference's loss (cross-entropy is 2-Lipschitz in the logits); contrastive-step
gradients no looser than 1.5x (global) / 3x (per tensor, floor 2.5e-2) the reference's own AMP deviation - the
gradient of the residual stream travels in bf16 here and in fp32 under HF autocast, which shows on the tensors that sum it
over many tokens (embedding tables: 2.6x at RoBERTa-large) while the global figure stays at 1.1x; probe gradients
cosine >= 0.999, rel-L2 <= 4e-2 (query weight / bias gradients, cancelling sums of dQ over tokens: 0.995 / 0.1).
"""
import pytest
import torch

from tests import realdims
from tests.util import cosine, load_golden, rel_l2, sub

pytestmark = pytest.mark.gpu


def _load_hf(enc, hf_model):
    enc._load_hf_state({k: v.detach().clone() for k, v in hf_model.state_dict().items()})


def _task(name, **kw):
    from dpr_scale_b200.task.dpr_task import DenseRetrieverTask
    kind, cfg, B, n, S, T = realdims.CASES[name]
    g = load_golden(f"realdims_{name}.npz")
    qm, cm = realdims.hf_models(kind, cfg)
    # the weights are rebuilt from the seed here: prove they are the ones the reference ran on
    assert torch.equal(realdims.checksums(qm), g["sum_q"]) and torch.equal(realdims.checksums(cm), g["sum_c"]), \
        "seeded weights differ from the ones the golden was generated with"
    task = DenseRetrieverTask(transform={}, datamodule=None, optim={}, shared_model=False, softmax_temperature=T,
                              model={"_target_": "dpr_scale_b200.models.hf_model.HFEncoder.from_config",
                                     "config": dict(cfg, model_type=kind), "dropout": 0.0}, **kw)
    task.trainer = None
    task.setup("fit")
    _load_hf(task.query_encoder, qm)
    _load_hf(task.context_encoder, cm)
    return task.cuda(), g, (qm, cm)


def _golden_batch(g):
    b = sub(g, "batch/")
    return {"query_ids": sub(b, "query_ids/"), "contexts_ids": sub(b, "contexts_ids/"),
            "pos_ctx_indices": b["pos_ctx_indices"], "ctx_mask": b["ctx_mask"].bool()}


def _check_step(name, loss_gate):
    kind, cfg, B, n, S, T = realdims.CASES[name]
    task, g, _ = _task(name)
    batch = _golden_batch(g)
    with torch.no_grad():
        q, c = task(batch["query_ids"], batch["contexts_ids"])
    eq, ec = rel_l2(q.cpu(), g["q_emb"]), rel_l2(c.cpu(), g["c_emb"])
    assert eq <= 1e-2 and ec <= 1e-2, (eq, ec)
    m = batch["ctx_mask"].repeat(q.shape[0], 1)
    logits = task.sim_score(q, c, m.cuda()).cpu() / T
    fin = torch.isfinite(g["logits"])
    assert torch.equal(torch.isfinite(logits), fin)
    dl = float((logits[fin] - g["logits"][fin]).abs().max())
    assert dl <= 1e-2 * float(g["logits"][fin].abs().max()), dl
    for e in (task.query_encoder, task.context_encoder):
        e.zero_grad()
    loss = task.training_step(batch, 0)
    amp_dev = abs(float(g["amp_loss"]) - float(g["loss"]))
    dloss = abs(float(loss.detach()) - float(g["loss"]))
    # (1) the scoring + cross-entropy kernel against the reference's formula (dpr_task.py:209-212) on OUR logits: tight
    own = logits.clone()
    own[~fin] = float("-inf")
    ce_own = float(torch.nn.functional.cross_entropy(own, batch["pos_ctx_indices"]))
    assert abs(float(loss.detach()) - ce_own) <= 2e-3, (float(loss.detach()), ce_own)
    # (2) against the reference's loss.  Softmax cross-entropy is 2-Lipschitz in max|dlogit|, and the logits of these
    # random-init models are raw 768/1024-wide dot products at temperature 1 (|logit| up to several hundred), so the
    # embedding gate above (1e-2 rel-L2) already implies a loss uncertainty well above SURVEY 8c's 5e-2: two
    # roundings of the same embedding accuracy (4.5e-3) gave 0.005 and 0.105 here.  Gate: 5e-2 OR the bound implied by
    # the measured logit error, whichever is larger.
    assert dloss <= max(loss_gate, 2.0 * dl), (float(loss.detach()), float(g["loss"]), "max|dlogit|", dl,
                                              "reference AMP deviation", amp_dev)
    loss.backward()
    torch.cuda.synchronize()
    names = realdims.sampled_grad_names(cfg)
    num = den = 0.0
    worst = (0.0, "")
    for side, enc in (("q", task.query_encoder), ("c", task.context_encoder)):
        params = dict(enc.named_parameters())
        for k in names:
            got = realdims.sample(k, params["transformer." + k].grad.detach().float().cpu())
            want, amp_rel = g[f"grad_{side}/{k}"], float(g[f"amp_rel_{side}/{k}"])
            if float(want.norm()) < 1e-9:
                continue
            rl = rel_l2(got, want)
            worst = max(worst, (rl / max(amp_rel, 2.5e-2), f"{side}/{k} rel {rl:.3g} amp {amp_rel:.3g}"))
            num += float(((got - want).double() ** 2).sum())
            den += float((want.double() ** 2).sum())
    sampled_rel = (num / den) ** 0.5
    print(f"{name}: emb rel {eq:.2e}/{ec:.2e}  max|dlogit| {dl:.3f}  |dloss| {dloss:.4f} (reference AMP {amp_dev:.4f})  "
          f"sampled-grad rel {sampled_rel:.3f} (reference AMP global {float(g['amp_global_rel']):.3f})  worst {worst}")
    assert worst[0] <= 3.0, worst
    assert sampled_rel <= 1.5 * float(g["amp_global_rel"]), (sampled_rel, float(g["amp_global_rel"]))
    return task, g


def _check_probe(task, g, name):
    kind, cfg, B, n, S, T = realdims.CASES[name]
    batch = _golden_batch(g)
    enc = task.context_encoder
    enc.zero_grad()
    rep = enc(batch["contexts_ids"])
    probe = torch.randn(rep.shape, generator=torch.Generator().manual_seed(3))
    (rep * probe.cuda()).sum().backward()
    torch.cuda.synchronize()
    params = dict(enc.named_parameters())
    worst = (1.0, 0.0, "")
    for k in realdims.sampled_grad_names(cfg):
        got = realdims.sample(k, params["transformer." + k].grad.detach().float().cpu())
        want = g[f"probe_c/{k}"]
        if float(want.norm()) < 1e-6 * float(g["probe_c/embeddings.LayerNorm.bias"].norm()):
            continue
        cs, rl = cosine(got, want), rel_l2(got, want)
        if cs < worst[0]:
            worst = (cs, rl, k)
        if k.endswith("self.query.bias") or k.endswith("self.query.weight"):
            # sums over all tokens of dQ (x^T dQ for the weight), whose terms largely cancel (the key-bias gradient is
            # identically zero for the same reason): at random init attention is near uniform, dQ is ~100x smaller than
            # the other gradients (|g| 0.3 vs 30 at layer 12) and the bf16 rounding of dS / dQ is visible here first.
            # Measured at RoBERTa-large S = 256, layer 12, for two roundings of the SAME arithmetic (LayerNorm with
            # scalar / packed fp32 instructions; every other tensor agrees to 1e-3 between the two): bias 0.9973 /
            # 0.9965, weight 0.9996 / 0.9970.
            assert cs >= 0.995 and rl <= 0.1, (k, cs, rl)
        else:
            assert cs >= 0.999 and rl <= 4e-2, (k, cs, rl)
    print(f"{name}: probe gradients worst cosine {worst}")


def test_bert_base_cfg1_training_step_matches_reference():
    # loss gate: SURVEY §8c's 5e-2 (the reference's own AMP deviates by 0.034 on this batch)
    task, g = _check_step("bert_base_cfg1", 5e-2)
    _check_probe(task, g, "bert_base_cfg1")


def test_bert_base_cfg1_all_gradients_match_live_oracle():
    """Every parameter gradient of the step (not only the committed sample) against the HF oracle on the host CPU."""
    from oracle import hf_path, task as otask
    name = "bert_base_cfg1"
    kind, cfg, B, n, S, T = realdims.CASES[name]
    task, g, (qm, cm) = _task(name)
    batch = _golden_batch(g)
    qe, ce = hf_path.CLSEncoder(None, model=qm), hf_path.CLSEncoder(None, model=cm)
    loss_ref, _ = otask.in_batch_loss(qe(batch["query_ids"]), ce(batch["contexts_ids"]), batch["ctx_mask"],
                                      batch["pos_ctx_indices"], T)
    assert abs(float(loss_ref) - float(g["loss"])) <= 1e-4      # the live oracle IS the reference (pinned by the golden)
    loss_ref.backward()
    for e in (task.query_encoder, task.context_encoder):
        e.zero_grad()
    task.training_step(batch, 0).backward()
    torch.cuda.synchronize()
    num = den = 0.0
    for enc, ref in ((task.query_encoder, qm), (task.context_encoder, cm)):
        rp = dict(ref.named_parameters())
        for k, p in enc.named_parameters():
            k = k[len("transformer."):]
            if rp[k].grad is None:
                assert p.grad is None or "pooler" in k
                continue
            num += float(((p.grad.detach().float().cpu() - rp[k].grad).double() ** 2).sum())
            den += float((rp[k].grad.double() ** 2).sum())
    rel = (num / den) ** 0.5
    print("bert_base_cfg1: global gradient rel-L2 over ALL parameters", rel, "reference AMP", float(g["amp_global_rel"]))
    assert rel <= 1.5 * float(g["amp_global_rel"])


def test_roberta_large_s256_training_step_matches_reference():
    # loss gate: the reference's own bf16 autocast is off by 0.27 on this batch (24 layers, |logit| ~ 190, 2 x 4
    # scores); the CUDA path must stay within 5e-2 + a quarter of that deviation.
    g = load_golden("realdims_roberta_large_s256.npz")
    gate = 5e-2 + 0.25 * abs(float(g["amp_loss"]) - float(g["loss"]))
    task, g = _check_step("roberta_large_s256", gate)
    _check_probe(task, g, "roberta_large_s256")


def test_fullsize_bench_batch_spot_check_against_oracle():
    """cfg 2's context batch at full size (1024 sequences x 128 tokens, BERT-base): 16 sequences spread over the batch
    (first / last / chunk borders) against the HF oracle, forward and backward.  The other 1008 sequences get a zero
    upstream gradient, so the parameter gradients must equal the oracle's gradients on the 16 alone."""
    from bench import BERT_BASE as BENCH_CFG, synth_batch
    from dpr_scale_b200.models.hf_model import HFEncoder
    kind, cfg = "bert", realdims.BERT_BASE
    qm, _ = realdims.hf_models(kind, cfg)
    enc = HFEncoder.from_config(dict(cfg, model_type="bert"), dropout=0.0)
    _load_hf(enc, qm)
    enc = enc.cuda().train()                      # dropout 0: train mode only selects the save-for-backward path
    S, N = 128, 1024
    tok = synth_batch(0, BENCH_CFG, 128, 7, S, pin=False)["contexts_ids"]
    am = tok["attention_mask"].clone()
    lens = torch.randint(S // 4, S + 1, (N,), generator=torch.Generator().manual_seed(5))
    am[1::2] = (torch.arange(S).unsqueeze(0) < lens[1::2].unsqueeze(1)).long()      # every other row padded
    tok = {"input_ids": tok["input_ids"] * am, "token_type_ids": tok["token_type_ids"], "attention_mask": am}
    pick = torch.tensor([0, 1, 2, 63, 64, 127, 128, 255, 256, 511, 512, 767, 1000, 1021, 1022, 1023])
    probe = torch.zeros(N, cfg["hidden_size"])
    probe[pick] = torch.randn(len(pick), cfg["hidden_size"], generator=torch.Generator().manual_seed(7))
    enc.zero_grad()
    rep = enc({k: v.cuda() for k, v in tok.items()})
    (rep * probe.cuda()).sum().backward()
    torch.cuda.synchronize()
    sub_tok = {k: v[pick] for k, v in tok.items()}
    ref = qm(**sub_tok)[0][:, 0, :]
    (ref * probe[pick]).sum().backward()
    got = rep.detach().cpu()[pick]
    per_row = ((got - ref.detach()).norm(dim=1) / ref.detach().norm(dim=1)).max()
    assert float(per_row) <= 1e-2, float(per_row)
    assert torch.isfinite(rep).all()
    num = den = 0.0
    rp = dict(qm.named_parameters())
    worst = (1.0, "")
    for k, p in enc.named_parameters():
        k = k[len("transformer."):]
        if rp[k].grad is None:
            continue
        a, b = p.grad.detach().float().cpu(), rp[k].grad
        num += float(((a - b).double() ** 2).sum())
        den += float((b.double() ** 2).sum())
        if float(b.norm()) > 1e-4 * float(rp["embeddings.LayerNorm.bias"].grad.norm()):
            worst = min(worst, (cosine(a, b), k))
    rel = (num / den) ** 0.5
    print(f"full-size spot check: worst pooled-row rel {float(per_row):.2e}, gradient rel-L2 {rel:.3e}, worst cosine {worst}")
    assert rel <= 4e-2 and worst[0] >= 0.998, (rel, worst)


TINY = dict(vocab_size=64, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
            max_position_embeddings=40)


def test_shared_model_equal_shapes_no_workspace_aliasing():
    """shared_model=True (reference default, dpr_task.py:25, :66-70): ONE encoder runs the query pass and the context
    pass; with equal shapes (num_negative = 0) both forwards are alive until backward.  Gradients of the shared
    parameters are the sum over both passes."""
    from dpr_scale_b200.task.dpr_task import DenseRetrieverTask
    from oracle import encoder as oenc, task as otask
    from tests.util import BERT_TINY_CFG
    g = load_golden("golden_1rank.npz")
    task = DenseRetrieverTask(transform={}, datamodule=None, optim={}, shared_model=True, softmax_temperature=8.0,
                              model={"_target_": "dpr_scale_b200.models.hf_model.HFEncoder.from_config",
                                     "config": TINY, "dropout": 0.0})
    task.trainer = None
    task.setup("fit")
    assert task.query_encoder is task.context_encoder
    task.query_encoder.load_state_dict(sub(g, "sd_q/"))
    task = task.cuda()
    ctx = sub(sub(g, "batch/"), "contexts_ids/")
    qtok = {k: v[:4] for k, v in ctx.items()}
    ctok = {k: v[4:8] for k, v in ctx.items()}
    assert qtok["input_ids"].shape == ctok["input_ids"].shape
    batch = {"query_ids": qtok, "contexts_ids": ctok, "pos_ctx_indices": torch.arange(4),
             "ctx_mask": torch.zeros(4, dtype=torch.bool)}
    sd = {k: v.clone().requires_grad_(True) for k, v in sub(g, "sd_q/").items()}
    pq = torch.randn(4, 128, generator=torch.Generator().manual_seed(11))
    pc = torch.randn(4, 128, generator=torch.Generator().manual_seed(12))
    rq, rc = oenc.encode(sd, BERT_TINY_CFG, qtok), oenc.encode(sd, BERT_TINY_CFG, ctok)
    want_loss, _ = otask.in_batch_loss(rq, rc, batch["ctx_mask"], batch["pos_ctx_indices"], 8.0)
    ((rq * pq).sum() + (rc * pc).sum()).backward()
    enc = task.query_encoder
    enc.zero_grad()
    q, c = task(batch["query_ids"], batch["contexts_ids"])
    assert enc._ws_pool.leased == 2                       # two live forwards, two distinct workspaces
    ((q * pq.cuda()).sum() + (c * pc.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert enc._ws_pool.leased == 0 and enc._pending_bwd == 0
    assert rel_l2(q.detach().cpu(), rq.detach()) <= 1e-2 and rel_l2(c.detach().cpu(), rc.detach()) <= 1e-2
    checked = 0
    for k, p in enc.named_parameters():
        r = sd[k].grad
        if r is None or float(r.norm()) < 1e-5:
            continue
        got = p.grad.detach().float().cpu()
        assert cosine(got, r) >= 0.9995 and rel_l2(got, r) <= 2e-2, (k, cosine(got, r), rel_l2(got, r))
        checked += 1
    assert checked >= 30
    enc.zero_grad()
    loss = task.training_step(batch, 0)
    assert abs(float(loss) - float(want_loss)) <= 5e-2
    loss.backward()
    torch.cuda.synchronize()
    assert enc._ws_pool.leased == 0


def test_projection_head_matches_oracle():
    """`projection_dim` (hf_model.py:26-34): Linear(H, p) + LayerNorm(p) after the CLS pooling, on the dprb kernels."""
    from dpr_scale_b200.models.hf_model import HFEncoder
    from oracle import encoder as oenc
    from tests.util import BERT_TINY_CFG
    g = load_golden("golden_1rank.npz")
    enc = HFEncoder.from_config(TINY, dropout=0.0, projection_dim=64)
    with torch.no_grad():
        gen = torch.Generator().manual_seed(21)
        enc.project[0].bias.copy_(0.05 * torch.randn(64, generator=gen))
        enc.project[1].weight.copy_(1 + 0.05 * torch.randn(64, generator=gen))
        enc.project[1].bias.copy_(0.05 * torch.randn(64, generator=gen))
    sdt = sub(g, "sd_c/")
    enc.transformer.load_state_dict({k[len("transformer."):]: v for k, v in sdt.items()})
    assert set(enc.state_dict()) == set(sdt) | {"project.0.weight", "project.0.bias", "project.1.weight", "project.1.bias"}
    proj = {k: v.detach().clone().requires_grad_(True) for k, v in enc.project.state_dict().items()}
    enc = enc.cuda()
    tok = sub(sub(g, "batch/"), "contexts_ids/")
    sd = {k: v.clone().requires_grad_(True) for k, v in sdt.items()}
    pooled = oenc.encode(sd, BERT_TINY_CFG, tok)
    ref = torch.nn.functional.layer_norm(pooled @ proj["0.weight"].T + proj["0.bias"], (64,), proj["1.weight"],
                                         proj["1.bias"], 1e-5)
    probe = torch.randn(ref.shape, generator=torch.Generator().manual_seed(22))
    (ref * probe).sum().backward()
    enc.zero_grad()
    rep = enc(tok)
    assert rep.shape == (8, 64) and rep.dtype == torch.float32
    assert rel_l2(rep.detach().cpu(), ref.detach()) <= 1e-2, rel_l2(rep.detach().cpu(), ref.detach())
    (rep * probe.cuda()).sum().backward()
    torch.cuda.synchronize()
    for k in ("0.weight", "0.bias", "1.weight", "1.bias"):
        got = dict(enc.project.named_parameters())[k].grad.detach().float().cpu()
        assert cosine(got, proj[k].grad) >= 0.999 and rel_l2(got, proj[k].grad) <= 3e-2, (k, cosine(got, proj[k].grad))
    for k, p in enc.named_parameters():                  # and the gradient flows on into the encoder body
        if not k.startswith("transformer.") or sd[k].grad is None or float(sd[k].grad.norm()) < 1e-5:
            continue
        got = p.grad.detach().float().cpu()
        assert cosine(got, sd[k].grad) >= 0.999, (k, cosine(got, sd[k].grad))
    with torch.no_grad():                                 # eval / no-grad path goes through the same kernels
        rep2 = enc.eval()(tok)
    assert torch.allclose(rep2, rep.detach(), atol=1e-5)
