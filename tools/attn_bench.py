#!/usr/bin/env python3
"""Time attention fwd/bwd: python tools/attn_bench.py S heads [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpr_scale_b200 import ops
S = int(sys.argv[1]); heads = int(sys.argv[2]); T = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
H = heads * 64; dev = "cuda"; bf = torch.bfloat16
def timeit(name, f, flops, iters=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:10s} S={S} heads={heads} T={T}: {ms*1e3:9.1f} us  {flops/ms/1e9:7.1f} TF/s", flush=True)
qkv = torch.randn(T, 3 * H, device=dev, dtype=bf)
ctx, lse = ops.attn_fwd(qkv, None, T // S, S, heads)
dctx = torch.randn(T, H, device=dev, dtype=bf)
fl = 4.0 * S * H * T
timeit("attn_fwd", lambda: ops.attn_fwd(qkv, None, T // S, S, heads), fl)
timeit("attn_bwd", lambda: ops.attn_bwd(qkv, None, ctx, lse, dctx, T // S, S, heads), 2.5 * fl)
