#!/usr/bin/env python3
"""Time LayerNorm fwd/bwd, colsum, attention fwd/bwd at the cfg-2 context shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpr_scale_b200 import ops
T, H, I, S, heads = 131072, 768, 3072, 128, 12
dev = "cuda"; bf = torch.bfloat16
def timeit(name, f, nbytes, iters=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} {ms*1e3:9.1f} us  {nbytes/ms/1e9:7.2f} TB/s", flush=True)
z = torch.randn(T, H, device=dev, dtype=bf); dy = torch.randn(T, H, device=dev, dtype=bf)
g = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
y, stats, _ = ops.ln_fwd(z, g, b, 1e-12)
dg, db_, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)
timeit("ln_fwd", lambda: ops.ln_fwd(z, g, b, 1e-12), T*H*4)
timeit("ln_bwd", lambda: ops.ln_bwd(dy, z, stats, g, dg, db_, dbias), T*H*6)
x = torch.randn(T, I, device=dev, dtype=bf); cs = torch.zeros(I, device=dev)
timeit("colsum [T,3072]", lambda: ops.colsum(x, cs), T*I*2)
qkv = torch.randn(T, 3*H, device=dev, dtype=bf)
ctx, lse = ops.attn_fwd(qkv, None, T//S, S, heads)
dctx = torch.randn(T, H, device=dev, dtype=bf)
timeit("attn_fwd", lambda: ops.attn_fwd(qkv, None, T//S, S, heads), T*H*2*4)
timeit("attn_bwd", lambda: ops.attn_bwd(qkv, None, ctx, lse, dctx, T//S, S, heads), T*H*2*8)
