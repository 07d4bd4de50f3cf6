"""Parity at BASELINE.json's FULL sizes (cfg 2: T = 131 072 context tokens, H = 768, I = 3072, S = 128) through
size-independent properties — the oracle cannot run these sizes in seconds, so each check reduces to a closed form:
GEMM checksums (1^T D and 1^T D 1 from operand row/column sums in fp64), attention with constant V (softmax rows sum
to 1 => output == V), LayerNorm row statistics, optimizer fixed point, encoder determinism."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
T, H, I, S, HEADS = 131072, 768, 3072, 128, 12
DEV = "cuda"
bf = torch.bfloat16


def test_gemm_fullsize_checksums():
    from dpr_scale_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(0)
    A = torch.randn(T, H, device=DEV, generator=g).to(bf)
    W = (torch.randn(I, H, device=DEV, generator=g) * 0.05).to(bf)
    bias = torch.randn(I, device=DEV, generator=g)
    D = torch.empty(T, I, device=DEV, dtype=bf)
    ops.gemm(A, W, D, T, I, H, H, H, I, False, False, ops.EPI_BIAS, bias)
    # column sums of D == (1^T A) W^T + T * bias
    want = A.double().sum(0) @ W.double().T + T * bias.double()
    got = D.double().sum(0)
    scale = float((A.double().abs().sum(0) @ W.double().abs().T).max())
    assert float((got - want).abs().max()) <= 2e-4 * scale, float((got - want).abs().max()) / scale
    # dgrad orientation (B read MN-major in place): dX = D W, checksum over columns
    dX = torch.empty(T, H, device=DEV, dtype=bf)
    ops.gemm(D, W, dX, T, H, I, I, H, H, False, True, ops.EPI_BIAS, None)
    want = D.double().sum(0) @ W.double()
    got = dX.double().sum(0)
    scale = float((D.double().abs().sum(0) @ W.double().abs()).max())
    assert float((got - want).abs().max()) <= 2e-4 * scale
    # wgrad (both operands MN-major, split-K fp32 atomics): dW = D^T A ; 1^T dW 1 = sum_t rowsum(D)_t rowsum(A)_t
    dW = torch.zeros(I, H, device=DEV)
    ops.gemm(D, A, dW, I, H, T, I, H, H, True, True, ops.EPI_F32_ATOMIC_ADD, None, splits=0)
    want = float((D.double().sum(1) * A.double().sum(1)).sum())
    ref_scale = float((D.double().abs().sum(1) * A.double().abs().sum(1)).sum())
    assert abs(float(dW.double().sum()) - want) <= 1e-5 * ref_scale
    # and a random 64x64 block of dW against the direct product
    blk = D[:, 1000:1064].double().T @ A[:, 300:364].double()
    l1 = D[:, 1000:1064].double().abs().T @ A[:, 300:364].double().abs()  # fp32 accumulation over K = 131 072 terms
    assert bool(((dW[1000:1064, 300:364].double() - blk).abs() <= 2e-5 * l1).all())


def test_attention_fullsize_constant_v_and_zero_grad_paths():
    from dpr_scale_b200 import ops
    nseq = T // S
    g = torch.Generator(device=DEV).manual_seed(1)
    qkv = torch.randn(T, 3 * H, device=DEV, generator=g).to(bf)
    qkv[:, 2 * H:] = 0.5  # V == const  =>  softmax(.) V == const whatever the scores are
    mask = torch.ones(nseq, S, dtype=torch.int32, device=DEV)
    mask[::3, 100:] = 0
    ctx, lse = ops.attn_fwd(qkv, mask, nseq, S, HEADS)
    assert torch.all((ctx.float() - 0.5).abs() <= 2 ** -8), float((ctx.float() - 0.5).abs().max())
    assert torch.isfinite(lse).all()
    # backward: with constant V, dP = dO.V^T is constant along keys => dS = P (dP - D) = 0 => dQ = dK = 0,
    # and dV_j = sum_i P_ij dO_i ; its column sums equal the column sums of dO over the (unmasked) queries
    dctx = torch.randn(T, H, device=DEV, generator=g).to(bf)
    dqkv = ops.attn_bwd(qkv, mask, ctx, lse, dctx, nseq, S, HEADS)
    assert float(dqkv[:, :2 * H].float().abs().max()) <= 2e-2
    got = dqkv[:, 2 * H:].float().view(nseq, S, H).sum(1)
    want = dctx.float().view(nseq, S, H).sum(1)
    assert float((got - want).abs().max()) <= 0.02 * float(want.abs().max()) + 0.05
    # masked keys receive no gradient at all
    assert float(dqkv.view(nseq, S, 3 * H)[::3, 100:, H:].float().abs().max()) == 0.0


def test_layernorm_fullsize_row_statistics():
    from dpr_scale_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(2)
    z = (torch.randn(T, H, device=DEV, generator=g) * 3 + 1).to(bf)
    y, stats, cls = ops.ln_fwd(z, torch.ones(H, device=DEV), torch.zeros(H, device=DEV), 1e-12, cls_stride=S)
    yf = y.float()
    assert float(yf.mean(1).abs().max()) <= 5e-3
    assert float((yf.var(1, unbiased=False) - 1).abs().max()) <= 2e-2
    assert cls.shape == (T // S, H) and float((cls - yf[::S]).abs().max()) <= 2 ** -7 * float(cls.abs().max())
    # backward is orthogonal to the constant vector and to xhat (gamma = 1): sum_c dz = 0, sum_c dz * xhat = 0
    dy = torch.randn(T, H, device=DEV, generator=g).to(bf)
    dg, db, dbias = (torch.zeros(H, device=DEV) for _ in range(3))
    dz = ops.ln_bwd(dy, z, stats, torch.ones(H, device=DEV), dg, db, dbias).float()
    assert float(dz.sum(1).abs().max()) <= 0.05 * float(dz.abs().sum(1).max())
    assert torch.allclose(db, dy.float().sum(0), rtol=1e-3, atol=1e-1)
    assert torch.allclose(dbias, dz.sum(0), rtol=1e-3, atol=1e-1)


def test_optimizer_fullsize_fixed_point_and_shadow():
    from dpr_scale_b200 import ops
    n = 108_891_648 + 3  # one BERT-base encoder arena (+ an odd tail)
    p = torch.randn(n, device=DEV)
    p0 = p.clone()
    gz = torch.zeros(n, device=DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    shadow = torch.empty(n, dtype=bf, device=DEV)
    ss = torch.zeros(1, device=DEV)
    ops.sumsq(gz, ss)
    ops.adamw_step(p, gz, m, v, shadow, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, ss, 2.0)
    assert float(ss) == 0.0 and torch.equal(p, p0)          # zero gradient, no weight decay => fixed point
    assert torch.equal(shadow, p0.to(bf))                    # shadow refreshed by the same kernel
    gr = torch.full((n,), 0.5, device=DEV)
    ss.zero_()
    ops.sumsq(gr, ss)
    assert abs(float(ss) - 0.25 * n) <= 1e-3 * 0.25 * n     # global-norm reduction at full size


def test_encoder_fullsize_forward_is_deterministic_and_well_scaled():
    from bench import BERT_BASE, synth_batch
    from dpr_scale_b200.models.hf_model import HFEncoder
    enc = HFEncoder.from_config(BERT_BASE, dropout=0.0).cuda().eval()
    tokens = {k: v.cuda() for k, v in synth_batch(0, BERT_BASE, 16, 7, S, pin=False)["contexts_ids"].items()}
    with torch.no_grad():
        a = enc(tokens)
        b = enc(tokens)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    # final LayerNorm with gamma = 1, beta = 0 => every pooled row has norm sqrt(H)
    assert float((a.norm(dim=1) - math.sqrt(H)).abs().max()) <= 0.05
