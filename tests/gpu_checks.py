"""Per-kernel parity checks: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Each check returns a dict of error figures and raises AssertionError when a stated tolerance is exceeded.
Used by tests/test_ops_gpu.py (pytest -m gpu) and tools/gpu_diag.py (one subprocess per check).

Tolerances (stated here, per the fp contract of BASELINE.json:north_star / SURVEY.md §8c):
  * bf16-output kernels: |err| <= 2^-7 * max|ref| + small atol   (one bf16 rounding of the result plus fp32
    accumulation-order noise; inputs are the identical bf16-rounded values on both sides)
  * fp32-output kernels (scoring/CE, optimizer, wgrad): rel <= 1e-4 unless noted.
"""
import math

import torch

from dpr_scale_b200 import ops
from oracle import encoder as oenc
from oracle import task as otask

DEV = "cuda"


def _bf(x):
    return x.to(torch.bfloat16)


def _close(name, got, ref, rtol_max, atol=0.0, out=None):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max())
    if out is not None:
        out[name + "_maxerr"] = err
        out[name + "_scale"] = scale
    assert err <= rtol_max * scale + atol, f"{name}: max err {err:.4e} > {rtol_max}*{scale:.4e}+{atol}"
    return err


# ------------------------------------------------------------------ GEMM
def check_gemm(M, N, K, a_mn=False, b_mn=False, epilogue=ops.EPI_BIAS, splits=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = _bf(torch.randn(M, K, generator=g))
    B = _bf(torch.randn(N, K, generator=g) * 0.5)
    bias = torch.randn(N, generator=g)
    aux = _bf(torch.randn(M, N, generator=g))
    ref = A.double() @ B.double().T
    res = {}
    Ad = (A.T.contiguous() if a_mn else A).to(DEV)
    Bd = (B.T.contiguous() if b_mn else B).to(DEV)
    lda = M if a_mn else K
    ldb = N if b_mn else K
    bias_d, aux_d = bias.to(DEV), aux.to(DEV)
    if epilogue in (ops.EPI_F32_ATOMIC_ADD, ops.EPI_F32_STORE):
        init = torch.randn(M, N, generator=g)
        out = init.clone().to(DEV)
        use_bias = epilogue == ops.EPI_F32_STORE
        ops.gemm(Ad, Bd, out, M, N, K, lda, ldb, N, a_mn, b_mn, epilogue, bias_d if use_bias else None, splits=splits)
        want = ref + (init.double() if epilogue == ops.EPI_F32_ATOMIC_ADD else bias.double())
        _close("gemm_f32", out, want, 2e-5 * math.sqrt(K), 1e-4, res)
        return res
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    out2 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV) if epilogue == ops.EPI_BIAS_GELU else None
    use_aux = epilogue in (ops.EPI_BIAS_RESIDUAL, ops.EPI_DGELU)
    use_bias = epilogue != ops.EPI_DGELU
    colsum = torch.ones(N, device=DEV) if epilogue != ops.EPI_BIAS_GELU else None
    ops.gemm(Ad, Bd, out, M, N, K, lda, ldb, N, a_mn, b_mn, epilogue, bias_d if use_bias else None,
             aux_d if use_aux else None, N if use_aux else 0, out2, colsum=colsum)
    torch.cuda.synchronize()
    if epilogue == ops.EPI_BIAS:
        want = ref + bias.double()
    elif epilogue == ops.EPI_BIAS_RESIDUAL:
        want = ref + bias.double() + aux.double()
    elif epilogue == ops.EPI_BIAS_GELU:
        pre = ref + bias.double()
        cdf = 0.5 * (1 + torch.erf(pre / math.sqrt(2)))
        pdf = torch.exp(-0.5 * pre * pre) / math.sqrt(2 * math.pi)
        # out2 = gelu'(pre) (what backward needs); fitted-CDF error <= 1.3e-4 + bf16 rounding
        _close("gemm_dgelu_saved", out2, cdf + pre * pdf, 2 ** -7, 1e-3, res)
        want = oenc.gelu_erf(pre)
    else:  # DGELU: plain multiply by the saved derivative
        want = ref * aux.double()
    _close("gemm_out", out, want, 2 ** -7, 1e-3, res)
    if colsum is not None:  # fused bias-gradient column sums of the (bf16-rounded) output
        _close("gemm_colsum", colsum, 1 + out.double().cpu().sum(0), 1e-5, 1e-3, res)
    return res


def check_gemm_f16_stream(M=384, N=768, K=512, seed=30):
    """The encoder's fp16 residual stream through the GEMM epilogue: bf16 operands, residual aux read as fp16, sum written
    as fp16 (DPRB_GEMM_{AUX,OUT}_F16); and the both-operands-fp16 form of the MMA (a mixed fp16 x bf16 pair is rejected)."""
    g = torch.Generator().manual_seed(seed)
    A = _bf(torch.randn(M, K, generator=g))
    B = _bf(torch.randn(N, K, generator=g) * 0.5)
    bias = torch.randn(N, generator=g)
    aux = (torch.randn(M, N, generator=g) * 3).half()
    res = {}
    out = torch.empty(M, N, dtype=torch.float16, device=DEV)
    ops.gemm(A.to(DEV), B.to(DEV), out, M, N, K, K, K, N, False, False,
             ops.EPI_BIAS_RESIDUAL | ops.GEMM_AUX_F16 | ops.GEMM_OUT_F16, bias.to(DEV), aux.to(DEV), N)
    want = A.double() @ B.double().T + bias.double() + aux.double()
    _close("gemm_f16_stream", out, want, 2 ** -10, 1e-3, res)          # fp16 output: 11 significand bits
    Ah, Bh = A.half(), B.half()
    out2 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(Ah.to(DEV), Bh.to(DEV), out2, M, N, K, K, K, N, False, False, ops.EPI_BIAS | ops.GEMM_A_F16 | ops.GEMM_B_F16,
             bias.to(DEV))
    _close("gemm_f16_ab", out2, Ah.double() @ Bh.double().T + bias.double(), 2 ** -7, 1e-3, res)
    try:
        ops.gemm(Ah.to(DEV), B.to(DEV), out2, M, N, K, K, K, N, False, False, ops.EPI_BIAS | ops.GEMM_A_F16, bias.to(DEV))
        raise AssertionError("mixed fp16 x bf16 operands must be rejected on the host")
    except Exception as e:  # noqa
        assert "fp16 x bf16" in str(e), e
    return res


def check_gemm_lean(M=384, N=512, K=256, seed=40):
    """The lean-activations epilogues: BIAS_GELU | SAVE_PRE (out2 = pre-activation), DGELU_PRE (acc * gelu'(pre) rebuilt
    from the saved pre-activation) and the elementwise gelu_from_pre - against erf-GELU and its exact derivative."""
    g = torch.Generator().manual_seed(seed)
    A = _bf(torch.randn(M, K, generator=g))
    B = _bf(torch.randn(N, K, generator=g) * 0.2)
    bias = torch.randn(N, generator=g)
    res = {}
    out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    pre_d = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A.to(DEV), B.to(DEV), out, M, N, K, K, K, N, False, False, ops.EPI_BIAS_GELU | ops.GEMM_SAVE_PRE, bias.to(DEV),
             out2=pre_d)
    pre = A.double() @ B.double().T + bias.double()
    _close("lean_pre", pre_d, pre, 2 ** -7, 1e-3, res)
    _close("lean_act", out, oenc.gelu_erf(pre), 2 ** -7, 1e-3, res)
    _close("lean_gelu_from_pre", ops.gelu_from_pre(pre_d), oenc.gelu_erf(pre_d.double().cpu()), 2 ** -7, 1e-3, res)
    # backward: dH[M, K2] = (dY W) * gelu'(pre) with B read MN-major; pre from the forward call above
    dY = _bf(torch.randn(M, K, generator=g))
    W = _bf(torch.randn(K, N, generator=g) * 0.2)          # [K_in = K, N_out = N] row-major == B(n, k) MN-major
    dH = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    cs = torch.zeros(N, device=DEV)
    ops.gemm(dY.to(DEV), W.to(DEV), dH, M, N, K, K, N, N, False, True, ops.EPI_DGELU_PRE, None, pre_d, N, colsum=cs)
    x = pre_d.double().cpu()
    cdf = 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    pdf = torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)
    want = (dY.double() @ W.double()) * (cdf + x * pdf)
    _close("lean_dgelu", dH, want, 2 ** -7, 2e-3, res)
    _close("lean_dgelu_colsum", cs, dH.double().cpu().sum(0), 1e-5, 1e-3, res)
    return res


# ------------------------------------------------------------------ LayerNorm / embeddings
def check_ln(T=777, H=768, cls_stride=0, seed=1, f16=False):
    """f16: z arrives in fp16 (written by a DPRB_GEMM_OUT_F16 epilogue) and the fp16 residual copy y_res is requested."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(T, H, generator=g) * 2 + 0.3
    z = z.half() if f16 else _bf(z)          # f16: the encoder's residual-stream format (z in / y out in fp16)
    gamma = 1 + 0.1 * torch.randn(H, generator=g)
    beta = 0.1 * torch.randn(H, generator=g)
    eps = 1e-12
    res = {}
    y_res = torch.empty(T, H, dtype=torch.float16, device=DEV) if f16 else None
    y, stats, cls = ops.ln_fwd(z.to(DEV), gamma.to(DEV), beta.to(DEV), eps, cls_stride, y_res)
    zr = z.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = oenc.layer_norm(zr, gr, br, eps)
    _close("ln_y", y, yr, 2 ** -7, 1e-3, res)
    if f16:
        _close("ln_y_res", y_res, yr, 2 ** -10, 1e-3, res)
    if cls_stride:
        _close("ln_cls", cls, yr[::cls_stride], 1e-5, 1e-5, res)
    # backward
    dy = _bf(torch.randn(T, H, generator=g))
    dgamma = torch.zeros(H, device=DEV)
    dbeta = torch.zeros(H, device=DEV)
    dbias = torch.zeros(H, device=DEV)
    if cls_stride:
        ncls = (T + cls_stride - 1) // cls_stride
        dy_cls = torch.randn(ncls, H, generator=g)
        dyr = torch.zeros(T, H)
        dyr[::cls_stride] = dy_cls
        dz = ops.ln_bwd(None, z.to(DEV), stats, gamma.to(DEV), dgamma, dbeta, dbias, dy_cls.to(DEV), cls_stride)
    else:
        dyr = dy.float()
        dz = ops.ln_bwd(dy.to(DEV), z.to(DEV), stats, gamma.to(DEV), dgamma, dbeta, dbias)
    yr.backward(dyr)
    _close("ln_dz", dz, zr.grad, 2 ** -7, 1e-3, res)
    _close("ln_dgamma", dgamma, gr.grad, 1e-4, 1e-3, res)
    _close("ln_dbeta", dbeta, br.grad, 1e-4, 1e-3, res)
    _close("ln_dbias", dbias, dz.float().cpu().sum(0), 1e-4, 1e-3, res)
    if not cls_stride:
        # hidden dropout between the Linear and this LayerNorm: second output dz * mask / (1-p), and the bias gradient
        # is the column sum of THAT (the Linear's own output gradient)
        p_drop, seed0, layer, site = 0.1, 11, 2, 3
        dg2, db2, dbias2 = (torch.zeros(H, device=DEV) for _ in range(3))
        dz2, dzm = ops.ln_bwd(dy.to(DEV), z.to(DEV), stats, gamma.to(DEV), dg2, db2, dbias2, None, 1, p_drop,
                              ops.dropout_site_seed(seed0, layer, site))
        keep = ops.dropout_mask(T, H, p_drop, seed0, layer, site).float()
        scale = 1.0 / (1.0 - round(p_drop * 65536) / 65536.0)
        assert torch.equal(dz2, dz), "ln_bwd: dz must not depend on the dropout site"
        _close("ln_drop_dzm", dzm, dz.float() * keep * scale, 2 ** -7, 1e-3, res)
        _close("ln_drop_dbias", dbias2, dzm.float().cpu().sum(0), 1e-4, 1e-3, res)
        _close("ln_drop_dgamma", dg2, gr.grad, 1e-4, 1e-3, res)
    return res


def check_embed(T=500, H=768, vocab=1000, max_pos=64, type_vocab=2, seed=2):
    g = torch.Generator().manual_seed(seed)
    word = torch.randn(vocab, H, generator=g) * 0.5
    pos = torch.randn(max_pos, H, generator=g) * 0.5
    typ = torch.randn(type_vocab, H, generator=g) * 0.5
    gamma = 1 + 0.1 * torch.randn(H, generator=g)
    beta = 0.1 * torch.randn(H, generator=g)
    ids = torch.randint(0, vocab, (T,), generator=g)
    ids[:50] = 7  # repeated ids -> atomic accumulation
    tts = torch.randint(0, type_vocab, (T,), generator=g)
    pids = torch.arange(T) % max_pos
    eps = 1e-12
    res = {}
    d = lambda t: t.to(DEV)
    y, stats = ops.embed_ln_fwd(d(ids), d(tts), d(pids), d(word), d(pos), d(typ), d(gamma), d(beta), eps)
    wr, pr, tr = (t.clone().requires_grad_(True) for t in (word, pos, typ))
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = oenc.layer_norm((wr[ids] + tr[tts]) + pr[pids], gr, br, eps)
    _close("emb_y", y, yr, 2 ** -7, 1e-3, res)
    dy = _bf(torch.randn(T, H, generator=g))
    dword, dpos, dtyp = torch.zeros_like(d(word)), torch.zeros_like(d(pos)), torch.zeros_like(d(typ))
    dgamma, dbeta = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.embed_ln_bwd(d(dy), d(ids), d(tts), d(pids), d(word), d(pos), d(typ), d(gamma), stats, dword, dpos, dtyp, dgamma, dbeta)
    yr.backward(dy.float())
    _close("emb_dword", dword, wr.grad, 1e-4, 1e-3, res)
    _close("emb_dpos", dpos, pr.grad, 1e-4, 1e-3, res)
    _close("emb_dtype", dtyp, tr.grad, 1e-4, 1e-3, res)
    _close("emb_dgamma", dgamma, gr.grad, 1e-4, 1e-3, res)
    _close("emb_dbeta", dbeta, br.grad, 1e-4, 1e-3, res)
    return res


def check_colsum(T=1000, N=2304, seed=3):
    g = torch.Generator().manual_seed(seed)
    x = _bf(torch.randn(T, N, generator=g))
    out = torch.ones(N, device=DEV)
    ops.colsum(x.to(DEV), out)
    res = {}
    _close("colsum", out, 1 + x.double().sum(0), 1e-5, 1e-3, res)
    return res


# ------------------------------------------------------------------ attention
def check_attention(nseq=3, S=128, heads=2, masked=True, seed=4, dropout=0.0):
    g = torch.Generator().manual_seed(seed)
    H = heads * 64
    T = nseq * S
    qkv = _bf(torch.randn(T, 3 * H, generator=g))
    am = torch.ones(nseq, S, dtype=torch.int32)
    if masked:
        for i in range(nseq):
            ln = int(torch.randint(max(1, S // 4), S + 1, (1,), generator=g))
            am[i, ln:] = 0
    res = {}
    dseed, site = 0x1234567, 0
    keepm = None
    if dropout > 0:  # attention-probability dropout: export the (never stored) mask and replay it in the oracle
        site = ops.dropout_site_seed(dseed, 3, 1)
        keepm = ops.dropout_mask(nseq * heads * S, S, dropout, dseed, 3, 1).view(nseq, heads, S, S).float().cpu()
        keepm = keepm / (1.0 - round(dropout * 65536) / 65536.0)
    ctx, lse = ops.attn_fwd(qkv.to(DEV), am.to(DEV) if masked else None, nseq, S, heads, True, dropout, site)
    qr = qkv.float().requires_grad_(True)
    x = qr.view(nseq, S, 3, heads, 64)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))  # nseq, heads, S, 64
    sc = q @ k.transpose(-1, -2) / 8.0
    if masked:
        sc = sc.masked_fill(am.view(nseq, 1, 1, S) == 0, float("-inf"))
    p = torch.softmax(sc, -1)
    if keepm is not None:
        p = p * keepm
    cr = (p @ v).transpose(1, 2).reshape(T, H)
    _close("attn_ctx", ctx, cr, 2 ** -7, 2e-3, res)
    _close("attn_lse", lse, torch.logsumexp(sc, -1), 1e-4, 1e-3, res)
    dctx = _bf(torch.randn(T, H, generator=g))
    dbias = torch.ones(3 * H, device=DEV)
    dqkv = ops.attn_bwd(qkv.to(DEV), am.to(DEV) if masked else None, ctx, lse, dctx.to(DEV), nseq, S, heads, dbias,
                        dropout, site)
    cr.backward(dctx.float())
    # P and dS are rounded to bf16 before the second matmuls (as in any flash-style kernel): 2^-6 headroom
    _close("attn_dqkv", dqkv, qr.grad, 2 ** -6, 4e-3, res)
    _close("attn_dbias", dbias, 1 + dqkv.double().cpu().sum(0), 1e-4, 1e-3, res)  # fused QKV bias gradient
    return res


# ------------------------------------------------------------------ scoring + CE
def check_score_ce(Q=37, C=250, d=768, inv_t=2.0, q0=8, nq=16, c0=40, nc=100, seed=5, pair=False):
    """Fused scoring + CE: the tensor-core single-pass path (d % 8 == 0; tiles recomputed in backward, no logits in HBM
    unless asked) AND the fp32 FFMA kernels, both against oracle/task.py (dpr_task.py:98-105, :197-212)."""
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(Q, d, generator=g)
    c = torch.randn(C, d, generator=g)
    mask = torch.rand(C, generator=g) < 0.1
    labels = torch.randint(0, C, (Q,), generator=g)
    mask[labels] = False
    res = {}
    d_ = lambda t: t.to(DEV)
    pm = None
    if pair:
        pm = torch.rand(Q, C, generator=g) < 0.2
        pm[torch.arange(Q), labels] = False
    qr, cr = q.clone().requires_grad_(True), c.clone().requires_grad_(True)
    full_mask = mask.unsqueeze(0).expand(Q, -1) if pm is None else (pm | mask.unsqueeze(0))
    logits_r = otask.sim_score(qr, cr, full_mask) * inv_t
    loss_r = torch.nn.functional.cross_entropy(logits_r, labels)
    loss_r.backward()
    fin = torch.isfinite(logits_r)
    pmd = None if pm is None else d_(pm.to(torch.uint8))

    def compare(tag, loss_sum, lse, logits, dq, dc):
        if logits is not None:
            assert torch.equal(torch.isfinite(logits.cpu()), fin)
            _close(tag + "logits", torch.where(fin, logits.cpu(), torch.zeros(())), torch.where(fin, logits_r.detach(), torch.zeros(())), 1e-5, 1e-3, res)
        _close(tag + "lse", lse, torch.logsumexp(logits_r.detach(), 1), 1e-5, 1e-3, res)
        res[tag + "loss_abs_err"] = abs(float(loss_sum) / Q - float(loss_r))
        assert res[tag + "loss_abs_err"] <= 1e-4 * max(1.0, abs(float(loss_r))), res
        _close(tag + "dq", dq, qr.grad[q0:q0 + nq], 1e-4, 1e-6, res)
        _close(tag + "dc", dc, cr.grad[c0:c0 + nc], 1e-4, 1e-6, res)

    # legacy fp32 FFMA kernels (stored logits)
    loss_sum, lse, logits = ops.score_ce_fwd_legacy(d_(q), d_(c), d_(mask.to(torch.uint8)), d_(labels), inv_t, True, pmd)
    dq, dc = ops.score_ce_bwd(d_(q), d_(c), logits, d_(labels), lse, 1.0, inv_t, q0, nq, c0, nc)
    compare("ffma_", loss_sum, lse, logits, dq, dc)
    if ops.score_tc_supported(Q, C, d):
        # training form: no logits, backward recomputes the local tiles
        loss_sum, lse, logits, ctx = ops.score_fwd(d_(q), d_(c), d_(mask.to(torch.uint8)), d_(labels), inv_t, False, pmd, (nq, nc))
        assert logits is None and ctx is not None
        dq, dc = ops.score_bwd(ctx, 1.0, inv_t, q0, nq, c0, nc)
        compare("tc_", loss_sum, lse, None, dq, dc)
        # evaluation form: logits requested; a second call on the same shapes (counters must have been reset)
        loss_sum2, lse2, logits2, _ = ops.score_fwd(d_(q), d_(c), d_(mask.to(torch.uint8)), d_(labels), inv_t, True, pmd)
        compare("tc2_", loss_sum2, lse2, logits2, dq, dc)
        assert torch.equal(lse2, lse)
    return res


# ------------------------------------------------------------------ optimizer
def check_adamw(n=100003, seed=6):
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n + 1, generator=g)[:n].clone()
    res = {}
    pd, md, vd = p.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    pr, mr, vr = p.clone(), torch.zeros(n), torch.zeros(n)
    for step in range(1, 4):
        gr = torch.randn(n, generator=g) * 3
        gd = gr.to(DEV)
        ss = torch.zeros(1, device=DEV)
        ops.sumsq(gd, ss)
        ops.adamw_step(pd, gd, md, vd, shadow, 1e-3, 0.9, 0.999, 1e-8, 0.01, step, 0.5, ss, 2.0)
        coef, total = otask.clip_coef([gr * 0.5], 2.0)
        res["norm_rel_err"] = abs(math.sqrt(float(ss)) * 0.5 - total) / total
        assert res["norm_rel_err"] < 1e-4
        otask.adamw_step(pr, gr * 0.5 * coef, mr, vr, step, 1e-3, weight_decay=0.01)
    _close("adam_p", pd, pr, 1e-5, 1e-6, res)
    _close("adam_m", md, mr, 1e-4, 1e-7, res)
    _close("adam_v", vd, vr, 1e-4, 1e-9, res)
    _close("adam_shadow", shadow, pr, 2 ** -8, 1e-6, res)
    return res


CHECKS = {
    "gemm_kk_bias_irregular": lambda: check_gemm(300, 520, 200),
    "gemm_kk_bias_big": lambda: check_gemm(1024, 768, 768),
    "gemm_kk_gelu": lambda: check_gemm(512, 1024, 256, epilogue=ops.EPI_BIAS_GELU),
    "gemm_kk_residual": lambda: check_gemm(384, 768, 512, epilogue=ops.EPI_BIAS_RESIDUAL),
    "gemm_kmn_dgelu": lambda: check_gemm(384, 512, 256, b_mn=True, epilogue=ops.EPI_DGELU),
    "gemm_kmn_bias": lambda: check_gemm(300, 520, 200, b_mn=True),
    "gemm_mnk_bias": lambda: check_gemm(256, 512, 320, a_mn=True),
    "gemm_mnmn_atomic_split": lambda: check_gemm(768, 768, 4096, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_ATOMIC_ADD, splits=0),
    "gemm_mnmn_atomic_irregular": lambda: check_gemm(200, 264, 1000, a_mn=True, b_mn=True, epilogue=ops.EPI_F32_ATOMIC_ADD, splits=3),
    "gemm_kk_f32_store": lambda: check_gemm(130, 260, 96, epilogue=ops.EPI_F32_STORE),
    "gemm_many_tiles": lambda: check_gemm(4096, 2304, 768),
    "ln_768": lambda: check_ln(777, 768),
    "ln_1024_cls": lambda: check_ln(512, 1024, cls_stride=64),
    "ln_128": lambda: check_ln(100, 128),
    "ln_768_f16": lambda: check_ln(777, 768, f16=True),
    "ln_1024_cls_f16": lambda: check_ln(512, 1024, cls_stride=64, f16=True),
    "gemm_lean_epilogues": lambda: check_gemm_lean(),
    "gemm_lean_epilogues_big": lambda: check_gemm_lean(1024, 3072, 768, seed=41),
    "gemm_f16_stream": lambda: check_gemm_f16_stream(),
    "gemm_f16_stream_big": lambda: check_gemm_f16_stream(1024, 768, 3072, seed=31),
    "embed": lambda: check_embed(),
    "colsum": lambda: check_colsum(),
    "attn_128_masked": lambda: check_attention(3, 128, 2, True),
    "attn_64_nomask": lambda: check_attention(2, 64, 3, False),
    "attn_100_masked": lambda: check_attention(2, 100, 2, True),
    "attn_256_masked": lambda: check_attention(2, 256, 1, True),
    "score_ce": lambda: check_score_ce(),
    "score_ce_small": lambda: check_score_ce(Q=8, C=16, d=128, inv_t=1.0, q0=0, nq=8, c0=0, nc=16),
    "adamw": lambda: check_adamw(),
}

# tcgen05 attention (S <= 128) extra shapes: many problems (persistent loop, barrier phases), heads=12
CHECKS["score_ce_8gpu_shape"] = lambda: check_score_ce(Q=1024, C=8192, d=768, inv_t=0.125, q0=256, nq=128, c0=2048,
                                                       nc=1024, seed=16)     # cfg 3: global 1024 x 8192 scores per rank
CHECKS["score_ce_pair_mask"] = lambda: check_score_ce(Q=130, C=300, d=128, inv_t=1.0, q0=1, nq=129, c0=0, nc=300, seed=17, pair=True)
CHECKS["score_ce_cfg4_shape"] = lambda: check_score_ce(Q=512, C=1024, d=1024, inv_t=1.0, q0=64, nq=64, c0=0, nc=1024, seed=18)
CHECKS["score_ce_odd_d"] = lambda: check_score_ce(Q=9, C=33, d=100, inv_t=1.0, q0=0, nq=9, c0=0, nc=33, seed=19)   # d % 8 != 0 -> FFMA only
CHECKS["score_ce_ragged_splits"] = lambda: check_score_ce(Q=70, C=1999, d=200, inv_t=0.5, q0=3, nq=60, c0=17, nc=1500,
                                                          seed=17)
CHECKS["attn_tc_many"] = lambda: check_attention(40, 128, 12, True, seed=9)
CHECKS["attn_tc_s64_many"] = lambda: check_attention(33, 64, 4, True, seed=10)
CHECKS["attn_tc_s37"] = lambda: check_attention(5, 37, 2, True, seed=11)
CHECKS["attn_tc2_s200"] = lambda: check_attention(3, 200, 2, True, seed=12)
CHECKS["attn_tc2_s256_many"] = lambda: check_attention(20, 256, 4, True, seed=13)
CHECKS["attn_tc_drop_s128"] = lambda: check_attention(4, 128, 2, True, seed=14, dropout=0.1)
CHECKS["attn_tc2_drop_s200"] = lambda: check_attention(3, 200, 2, True, seed=15, dropout=0.1)


# ------------------------------------------------------------------ retrieval
def check_search(Q=100, N=5000, d=768, k=100, bf16=False, seed=20, mode="random", offset=0):
    """dprb_search_topk vs oracle/retrieval.py (restating run_retrieval_pytorch.py:141-176) on float64 scores.

    Tolerance: the kernel ranks by fp32-accumulated products of the identical 16-bit operands, so a returned score
    may differ from the float64 score of the same row by tol = 1e-5 * max_row(|q|.|c|); ids must agree with the
    oracle wherever scores are separated by more than 2 tol, and the returned set must be a valid top-k up to 2 tol.
    """
    import numpy as np
    from oracle import retrieval as R
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(Q, d, generator=g)
    if mode == "random":
        c = torch.randn(N, d, generator=g)
    elif mode == "dup":          # every row appears twice -> exact score ties, resolved towards the lower id
        h = torch.randn((N + 1) // 2, d, generator=g)
        c = torch.cat([h, h], 0)[:N]
    elif mode == "ascending":    # scores grow with the row id for every query: worst case for the running threshold
        v = torch.randn(d, generator=g)
        q = q * 0.05 + v
        c = v[None, :] * (0.25 + torch.arange(N, dtype=torch.float32)[:, None] / N)
    elif mode == "constant":     # every row identical: all scores tie, the answer is rows 0..k-1 for every query
        c = torch.randn(1, d, generator=g).expand(N, d).contiguous()
    else:
        raise ValueError(mode)
    dt = torch.bfloat16 if bf16 else torch.float16
    q16, c16 = q.to(dt), c.to(dt)
    s, i = ops.search_topk(q16.to(DEV), c16.to(DEV), k, index_offset=offset)
    torch.cuda.synchronize()
    s, i = s.cpu().numpy().astype(np.float64), i.cpu().numpy() - offset
    qe, ce = q16.double().numpy(), c16.double().numpy()
    exact = qe @ ce.T
    tol = 1e-5 * float((np.abs(qe) @ np.abs(ce).T).max())
    out = {"tol": tol}
    assert s.shape == (Q, k) and i.shape == (Q, k)
    assert (i >= 0).all() and (i < N).all(), "row ids out of range"
    assert all(len(set(r)) == k for r in i), "duplicate row ids"
    ds = np.diff(s, axis=1)
    assert (ds <= 0).all(), "scores not descending"
    assert (np.diff(i, axis=1)[ds == 0] > 0).all(), "equal scores not ordered by ascending row id"
    got = np.take_along_axis(exact, i, axis=1)
    out["score_maxerr"] = float(np.abs(got - s).max())
    assert out["score_maxerr"] <= tol, f"score error {out['score_maxerr']:.3e} > {tol:.3e}"
    os_, oi = R.topk_desc(exact, min(k + 1, N))
    kth = os_[:, k - 1]
    assert (got >= kth[:, None] - 2 * tol).all(), "returned a row that is not in the top-k"
    mism = 0
    for r in range(Q):
        must = oi[r, :k][os_[r, :k] > kth[r] + 2 * tol]
        assert set(must) <= set(i[r]), f"query {r}: missing rows with score above the k-th"
        e = os_[r]
        for j in range(k):
            lo = e[j] - e[j + 1] if j + 1 < len(e) else np.inf
            hi = e[j - 1] - e[j] if j > 0 else np.inf
            if lo > 2 * tol and hi > 2 * tol:
                assert i[r, j] == oi[r, j], f"query {r} rank {j}: {i[r, j]} vs oracle {oi[r, j]}"
            else:
                mism += int(i[r, j] != oi[r, j])
    out["near_tie_swaps"] = mism
    return out


def check_topk_merge(Q=37, total=300, k=100, seed=30):
    """dprb_topk_merge vs the restated shard merge (run_retrieval_pytorch.py:272-277: topk over the concatenated shard
    results + gather); bit-exact (fp32 compare / select only), ties towards the earlier position."""
    import numpy as np
    from oracle import retrieval as R
    g = torch.Generator().manual_seed(seed)
    s = torch.randn(Q, total, generator=g)
    nt = min(s[:, ::7].shape[1], s[:, 3::7].shape[1])
    s[:, 0:7 * nt:7] = s[:, 3:3 + 7 * nt:7]                  # inject exact ties
    idx = torch.randint(0, 2 ** 40, (Q, total), generator=g, dtype=torch.int64)
    ms, mi = ops.topk_merge(s.to(DEV), idx.to(DEV), k)
    torch.cuda.synchronize()
    rs, order = R.topk_desc(s.numpy().astype(np.float64), k)
    ri = np.take_along_axis(idx.numpy(), order, axis=1)
    assert np.array_equal(ms.cpu().numpy().astype(np.float64), rs), "merged scores differ"
    assert np.array_equal(mi.cpu().numpy(), ri), "merged ids differ"
    return {"exact": True}


CHECKS["search_fp16_100x5000"] = lambda: check_search(100, 5000, 768, 100)
CHECKS["search_small_ragged"] = lambda: check_search(7, 1000, 200, 10, seed=21)
CHECKS["search_k_equals_n_region"] = lambda: check_search(5, 300, 64, 256, seed=22)
CHECKS["search_one_query_k1"] = lambda: check_search(1, 4097, 128, 1, seed=23)
CHECKS["search_bf16_300x70000"] = lambda: check_search(300, 70000, 1024, 256, bf16=True, seed=24, offset=1000000)
CHECKS["search_many_query_tiles"] = lambda: check_search(1100, 30000, 256, 100, seed=25)
CHECKS["search_k1000"] = lambda: check_search(40, 60000, 128, 1000, seed=26)
CHECKS["search_duplicate_rows"] = lambda: check_search(33, 9000, 128, 50, seed=27, mode="dup")
CHECKS["search_ascending_scores"] = lambda: check_search(64, 20000, 64, 100, seed=28, mode="ascending")
CHECKS["search_all_equal"] = lambda: check_search(40, 5000, 64, 100, seed=29, mode="constant")
CHECKS["topk_merge"] = lambda: check_topk_merge()
CHECKS["topk_merge_large"] = lambda: check_topk_merge(Q=5, total=40000, k=1000, seed=31)
