"""Tensor-level wrappers over the C ABI (raw device pointers + the current CUDA stream).

PyTorch is plumbing here: device memory, streams, autograd bookkeeping.  Every function launches
hand-written sm_100a kernels from libdprb.so; nothing falls back to torch math.
"""
import torch

from . import _lib
from ._lib import check

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_F32_ATOMIC_ADD, EPI_F32_STORE, EPI_DGELU_PRE = range(7)
GEMM_SAVE_PRE = 0x1000
# OR-ed into the epilogue: that operand holds fp16 instead of bf16 (include/dprb.h DPRB_GEMM_*_F16)
GEMM_A_F16, GEMM_B_F16, GEMM_AUX_F16, GEMM_OUT_F16 = 0x100, 0x200, 0x400, 0x800

def launch_count():
    """Kernels launched by libdprb.so in this process so far (counted inside the C launchers)."""
    return int(_lib.load().dprb_launch_count())


def _count(n=1):  # kept as a no-op hook: launches are counted in C (dprb_launch_count), not estimated here
    pass


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "dprb ops need CUDA tensors (no CPU fallback)"
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def gemm(a, b, out, M, N, K, lda, ldb, ldd, a_mn=False, b_mn=False, epilogue=EPI_BIAS, bias=None, aux=None,
         ld_aux=0, out2=None, alpha=1.0, splits=1, colsum=None, dropout_p=0.0, drop_seed=0):
    lib = _lib.load()
    check(lib.dprb_gemm_bf16(_ptr(a), _ptr(b), _ptr(out), M, N, K, lda, ldb, ldd, int(a_mn), int(b_mn), epilogue,
                             _ptr(bias), _ptr(aux), ld_aux, _ptr(out2), float(alpha), splits, _ptr(colsum),
                             float(dropout_p), int(drop_seed), _stream()),
          "dprb_gemm_bf16")
    _count()
    return out


def linear_fwd(x, w, bias=None, epilogue=EPI_BIAS, aux=None, out2=None):
    """y[T,N] = epi(x[T,K] @ w[N,K]^T + bias); x, w bf16 contiguous; bias fp32."""
    T, K = x.shape
    N = w.shape[0]
    y = torch.empty(T, N, dtype=torch.bfloat16, device=x.device)
    gemm(x, w, y, T, N, K, K, K, N, False, False, epilogue, bias, aux, N if aux is not None else 0, out2)
    return y


def embed_ln_fwd(ids, type_ids, pos_ids, word, pos, typ, gamma, beta, eps, dropout_p=0.0, seed=0, y_res=None):
    """y_res: optional fp16 [T, H] tensor that receives a second copy of the output (the residual-stream copy)."""
    T = ids.numel()
    H = word.shape[1]
    y = torch.empty(T, H, dtype=torch.bfloat16, device=word.device)
    stats = torch.empty(T, 2, dtype=torch.float32, device=word.device)
    check(_lib.load().dprb_embed_ln_fwd(_ptr(ids), _ptr(type_ids), _ptr(pos_ids), _ptr(word), _ptr(pos), _ptr(typ),
                                        _ptr(gamma), _ptr(beta), _ptr(y), _ptr(stats), T, H, word.shape[0],
                                        pos.shape[0], typ.shape[0], float(eps), float(dropout_p), int(seed), _ptr(y_res),
                                        _stream()), "dprb_embed_ln_fwd")
    _count()
    return y, stats


def embed_ln_bwd(dy, ids, type_ids, pos_ids, word, pos, typ, gamma, stats, dword, dpos, dtyp, dgamma, dbeta,
                 dropout_p=0.0, seed=0):
    T = ids.numel()
    H = word.shape[1]
    check(_lib.load().dprb_embed_ln_bwd(_ptr(dy), _ptr(ids), _ptr(type_ids), _ptr(pos_ids), _ptr(word), _ptr(pos),
                                        _ptr(typ), _ptr(gamma), _ptr(stats), _ptr(dword), _ptr(dpos), _ptr(dtyp),
                                        _ptr(dgamma), _ptr(dbeta), T, H, float(dropout_p), int(seed), _stream()),
          "dprb_embed_ln_bwd")
    _count()


def ln_fwd(z, gamma, beta, eps, cls_stride=0, y_res=None):
    """z: bf16, or fp16 (the encoder's residual-stream sums); y: bf16; y_res: optional fp16 copy of y."""
    T, H = z.shape
    y = torch.empty(T, H, dtype=torch.bfloat16, device=z.device)
    stats = torch.empty(T, 2, dtype=torch.float32, device=z.device)
    cls = None
    if cls_stride:
        cls = torch.empty((T + cls_stride - 1) // cls_stride, H, dtype=torch.float32, device=z.device)
    check(_lib.load().dprb_ln_fwd(_ptr(z), _ptr(gamma), _ptr(beta), _ptr(y), _ptr(stats), _ptr(cls),
                                  cls_stride if cls_stride else 1, T, H, float(eps), int(z.dtype == torch.float16),
                                  _ptr(y_res), _stream()), "dprb_ln_fwd")
    _count()
    return y, stats, cls


def ln_bwd(dy, z, stats, gamma, dgamma, dbeta, dbias=None, dy_cls=None, cls_stride=1, dropout_p=0.0, site_seed=0):
    T, H = z.shape
    dz = torch.empty(z.shape, dtype=torch.bfloat16, device=z.device)      # gradients are bf16 whatever z holds
    dzm = torch.empty_like(dz) if dropout_p > 0 else None
    check(_lib.load().dprb_ln_bwd(_ptr(dy), _ptr(dy_cls), cls_stride, _ptr(z), _ptr(stats), _ptr(gamma), _ptr(dz),
                                  _ptr(dgamma), _ptr(dbeta), _ptr(dbias), T, H, _ptr(dzm), float(dropout_p),
                                  int(site_seed), int(z.dtype == torch.float16), _stream()), "dprb_ln_bwd")
    _count()
    return (dz, dzm) if dropout_p > 0 else dz


def dropout_site_seed(seed, layer, site):
    return int(_lib.load().dprb_dropout_site_seed(int(seed), layer, site))


def dropout_mask(rows, cols, p, seed, layer, site, device="cuda"):
    """keep mask (uint8 [rows, cols]) of one dropout site — test aid."""
    out = torch.empty(rows, cols, dtype=torch.uint8, device=device)
    check(_lib.load().dprb_dropout_mask(_ptr(out), rows, cols, float(p), int(seed), layer, site, _stream()),
          "dprb_dropout_mask")
    return out


def gelu_from_pre(pre):
    out = torch.empty_like(pre)
    check(_lib.load().dprb_gelu_from_pre(_ptr(pre), _ptr(out), pre.numel(), _stream()), "dprb_gelu_from_pre")
    return out


def colsum(x, out):
    T, N = x.shape
    check(_lib.load().dprb_colsum_bf16(_ptr(x), x.stride(0), _ptr(out), T, N, _stream()), "dprb_colsum_bf16")
    _count()
    return out


def attn_fwd(qkv, attn_mask, nseq, S, heads, need_lse=True, dropout_p=0.0, site_seed=0):
    T = nseq * S
    H = heads * 64
    ctx = torch.empty(T, H, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(nseq, heads, S, dtype=torch.float32, device=qkv.device) if need_lse else None
    check(_lib.load().dprb_attn_fwd(_ptr(qkv), _ptr(attn_mask), _ptr(ctx), _ptr(lse), nseq, S, heads, float(dropout_p),
                                    int(site_seed), _stream()), "dprb_attn_fwd")
    _count()
    return ctx, lse


def attn_bwd(qkv, attn_mask, ctx, lse, dctx, nseq, S, heads, dbias=None, dropout_p=0.0, site_seed=0):
    dqkv = torch.empty_like(qkv)
    check(_lib.load().dprb_attn_bwd(_ptr(qkv), _ptr(attn_mask), _ptr(ctx), _ptr(lse), _ptr(dctx), _ptr(dqkv),
                                    _ptr(dbias), nseq, S, heads, float(dropout_p), int(site_seed), _stream()),
          "dprb_attn_bwd")
    _count()
    return dqkv


import os as _os


class ScoreCtx:
    """What the tensor-core scoring forward leaves for its backward: the workspace holding the exact bf16 operand splits
    (and room for the recomputed W tiles), plus the masks / labels / lse the recomputation needs."""

    __slots__ = ("ws", "col_mask", "pair_mask", "labels", "lse", "shape", "local")

    def __init__(self, ws, col_mask, pair_mask, labels, lse, shape, local):
        self.ws, self.col_mask, self.pair_mask, self.labels, self.lse = ws, col_mask, pair_mask, labels, lse
        self.shape, self.local = shape, local


def score_tc_supported(Q, C, d):
    return bool(_lib.load().dprb_score_tc_supported(Q, C, d)) and not _os.environ.get("DPRB_SCORE_LEGACY")


def score_fwd(q, c, col_mask, labels, inv_temperature, want_logits=False, pair_mask=None, local=None):
    """Fused scoring + CE forward.  Returns (loss_sum[1], lse[Q], logits or None, ctx): `ctx` (ScoreCtx) is what
    score_bwd needs on the tensor-core path, or None when the shape falls back to the FFMA kernels (d % 8 != 0), in
    which case logits are always produced (that backward reads them).  local = (nq, nc) sizes backward's W tiles."""
    Q, d = q.shape
    C = c.shape[0]
    lib = _lib.load()
    lse = torch.empty(Q, dtype=torch.float32, device=q.device)
    loss_sum = torch.zeros(1, dtype=torch.float32, device=q.device)
    if not score_tc_supported(Q, C, d):
        logits = torch.empty(Q, C, dtype=torch.float32, device=q.device)
        check(lib.dprb_score_ce_fwd(_ptr(q), _ptr(c), _ptr(col_mask), _ptr(pair_mask), _ptr(labels),
                                    float(inv_temperature), _ptr(lse), _ptr(loss_sum), _ptr(logits), Q, C, d, _stream()),
              "dprb_score_ce_fwd")
        return loss_sum, lse, logits, None
    nq, nc = local if local is not None else (0, 0)
    nbytes = lib.dprb_score_tc_workspace_bytes(Q, C, d, nq, nc)
    ws = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=q.device)
    off = (-ws.data_ptr()) % 256
    logits = torch.empty(Q, C, dtype=torch.float32, device=q.device) if want_logits else None
    check(lib.dprb_score_tc_fwd(_ptr(q), _ptr(c), _ptr(col_mask), _ptr(pair_mask), _ptr(labels), float(inv_temperature),
                                _ptr(lse), _ptr(loss_sum), _ptr(logits), Q, C, d, nq, nc, ws.data_ptr() + off,
                                ws.numel() - off, _stream()), "dprb_score_tc_fwd")
    return loss_sum, lse, logits, ScoreCtx(ws, col_mask, pair_mask, labels, lse, (Q, C, d), (nq, nc))


def score_bwd(ctx, grad_scale, inv_temperature, q0, nq, c0, nc):
    """dq[nq, d], dc[nc, d] of mean-over-Q CE for the rank-local rows / columns; tiles are recomputed (ScoreCtx)."""
    Q, C, d = ctx.shape
    assert (nq, nc) == tuple(ctx.local), "score_fwd must be told the local (nq, nc) that backward asks for"
    dq = torch.empty(nq, d, dtype=torch.float32, device=ctx.ws.device)
    dc = torch.empty(nc, d, dtype=torch.float32, device=ctx.ws.device)
    off = (-ctx.ws.data_ptr()) % 256
    check(_lib.load().dprb_score_tc_bwd(_ptr(ctx.col_mask), _ptr(ctx.pair_mask), _ptr(ctx.labels), _ptr(ctx.lse),
                                        float(grad_scale), float(inv_temperature), _ptr(dq), _ptr(dc), Q, C, d, q0,
                                        nq, c0, nc, ctx.ws.data_ptr() + off, ctx.ws.numel() - off, _stream()),
          "dprb_score_tc_bwd")
    return dq, dc


def score_ce_fwd(q, c, col_mask, labels, inv_temperature, want_logits=True, pair_mask=None):
    """(loss_sum, lse, logits) - the forward-only form used by sim_score / evaluation."""
    loss_sum, lse, logits, _ = score_fwd(q, c, col_mask, labels, inv_temperature, want_logits, pair_mask)
    return loss_sum, lse, logits


def score_ce_fwd_legacy(q, c, col_mask, labels, inv_temperature, want_logits=True, pair_mask=None):
    """The fp32 FFMA kernels (any d): kept for d % 8 != 0 and as the A/B reference of the tensor-core path."""
    Q, d = q.shape
    C = c.shape[0]
    lse = torch.empty(Q, dtype=torch.float32, device=q.device)
    loss_sum = torch.zeros(1, dtype=torch.float32, device=q.device)
    logits = torch.empty(Q, C, dtype=torch.float32, device=q.device) if want_logits else None
    check(_lib.load().dprb_score_ce_fwd(_ptr(q), _ptr(c), _ptr(col_mask), _ptr(pair_mask), _ptr(labels), float(inv_temperature),
                                        _ptr(lse), _ptr(loss_sum), _ptr(logits), Q, C, d, _stream()),
          "dprb_score_ce_fwd")
    return loss_sum, lse, logits


def score_ce_bwd(q, c, logits, labels, lse, grad_scale, inv_temperature, q0, nq, c0, nc):
    Q, d = q.shape
    C = c.shape[0]
    dq = torch.empty(nq, d, dtype=torch.float32, device=q.device)
    dc = torch.empty(nc, d, dtype=torch.float32, device=q.device)
    check(_lib.load().dprb_score_ce_bwd(_ptr(q), _ptr(c), _ptr(logits), _ptr(labels), _ptr(lse), float(grad_scale),
                                        float(inv_temperature), _ptr(dq), _ptr(dc), Q, C, d, q0, nq, c0, nc,
                                        _stream()), "dprb_score_ce_bwd")
    return dq, dc


def sumsq(g, out):
    check(_lib.load().dprb_sumsq_f32(_ptr(g), g.numel(), _ptr(out), _stream()), "dprb_sumsq_f32")
    _count()
    return out


def adamw_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, sumsq_buf=None,
               max_norm=0.0):
    check(_lib.load().dprb_adamw_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(shadow), p.numel(), float(lr),
                                      float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                                      float(grad_scale), _ptr(sumsq_buf), float(max_norm), _stream()),
          "dprb_adamw_step")
    _count()


def cast_f32_bf16(src, dst):
    check(_lib.load().dprb_cast_f32_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "dprb_cast_f32_bf16")
    _count()
    return dst


def cast_bf16_f32(src, dst):
    check(_lib.load().dprb_cast_bf16_f32(_ptr(src), _ptr(dst), src.numel(), _stream()), "dprb_cast_bf16_f32")
    return dst


_SEARCH_WS = {}


def _search_ws(nbytes, device):
    """Caller-owned workspace, cached per device and grown on demand (queues are reused across calls)."""
    buf = _SEARCH_WS.get(device)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _SEARCH_WS[device] = buf
    return buf


def search_topk(queries, corpus, k, index_offset=0, reference_ranking=False):
    """Fused inner-product search + top-k: queries [Q, d], corpus [N, d] (both fp16 or both bf16, contiguous)
    -> (scores fp32 [Q, k] descending, row ids int64 [Q, k]); never materialises the [Q, N] score matrix.
    reference_ranking: order by the fp16-rounded score, as the reference's topk over its fp16 einsum does."""
    assert queries.dtype == corpus.dtype and queries.dtype in (torch.float16, torch.bfloat16)
    assert queries.is_contiguous() and corpus.is_contiguous() and queries.shape[1] == corpus.shape[1]
    lib = _lib.load()
    Q, d = queries.shape
    N = corpus.shape[0]
    ws = _search_ws(lib.dprb_search_workspace_bytes(Q, int(k)), queries.device)
    scores = torch.empty(Q, k, dtype=torch.float32, device=queries.device)
    index = torch.empty(Q, k, dtype=torch.int64, device=queries.device)
    check(lib.dprb_search_topk(_ptr(queries), _ptr(corpus),
                               (1 if queries.dtype == torch.bfloat16 else 0) | (0x100 if reference_ranking else 0), Q, N, d,
                               int(k), int(index_offset), _ptr(scores), _ptr(index), _ptr(ws), ws.numel(),
                               _stream()), "dprb_search_topk")
    _count(2 * ((Q + 1023) // 1024))
    return scores, index


def topk_merge(scores, index, k):
    """k best of each row of scores [Q, total] fp32 with their index [Q, total] int64 entries (shard merge)."""
    assert scores.dtype == torch.float32 and index.dtype == torch.int64 and scores.shape == index.shape
    assert scores.is_contiguous() and index.is_contiguous()
    lib = _lib.load()
    Q, total = scores.shape
    ws = torch.empty(int(lib.dprb_topk_merge_workspace_bytes(Q, total)), dtype=torch.uint8, device=scores.device)
    out_s = torch.empty(Q, k, dtype=torch.float32, device=scores.device)
    out_i = torch.empty(Q, k, dtype=torch.int64, device=scores.device)
    check(lib.dprb_topk_merge(_ptr(scores), _ptr(index), Q, total, int(k), _ptr(out_s), _ptr(out_i), _ptr(ws),
                              ws.numel(), _stream()), "dprb_topk_merge")
    _count(2)
    return out_s, out_i
