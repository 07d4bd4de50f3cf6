#!/usr/bin/env python3
"""LayerNorm forward / backward alone at the encoder's shapes: microseconds and HBM bytes per second, packed-fp32
full-width kernels vs the generic ones (DPRB_LN_GENERIC=1), plus the largest difference between their outputs.

  python tools/ln_bench.py            # BERT-base ctx batch (131 072 x 768) and RoBERTa-large cfg 4 (278 528 x 1024)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpr_scale_b200 import ops

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def timed(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3


def run(T, H):
    g = torch.Generator(device=dev).manual_seed(0)
    z = (torch.randn(T, H, device=dev, generator=g) * 2 + 0.3).half()
    dy = torch.randn(T, H, device=dev, generator=g).bfloat16()
    gamma = 1 + 0.1 * torch.randn(H, device=dev, generator=g)
    beta = 0.1 * torch.randn(H, device=dev, generator=g)
    yres = torch.empty(T, H, dtype=torch.float16, device=dev)
    seed = ops.dropout_site_seed(7, 3, 2)
    out = {"T": T, "H": H}
    keep = {}
    for mode in ("full", "generic"):
        os.environ["DPRB_LN_GENERIC"] = "1" if mode == "generic" else "0"
        y, stats, _ = ops.ln_fwd(z, gamma, beta, 1e-12, 0, yres)
        dg, db, dbias = (torch.zeros(H, device=dev) for _ in range(3))
        dz, dzm = ops.ln_bwd(dy, z, stats, gamma, dg, db, dbias, None, 1, 0.1, seed)
        keep[mode] = [t.float().clone() for t in (y, yres, stats, dz, dzm, dg, db, dbias)]
        t_f = timed(lambda: ops.ln_fwd(z, gamma, beta, 1e-12, 0, yres))
        t_b = timed(lambda: ops.ln_bwd(dy, z, stats, gamma, dg, db, dbias, None, 1, 0.1, seed))
        out[mode] = {"fwd_us": round(t_f, 1), "fwd_TBs": round(T * H * 6 / t_f / 1e6, 2),
                     "bwd_us": round(t_b, 1), "bwd_TBs": round(T * H * 8 / t_b / 1e6, 2)}
    os.environ["DPRB_LN_GENERIC"] = "0"
    names = ("y", "y_res", "stats", "dz", "dzm", "dgamma", "dbeta", "dbias")
    out["max_abs_diff_full_vs_generic"] = {n: float((a - b).abs().max()) for n, a, b in zip(names, keep["full"], keep["generic"])}
    out["rel_l2_diff"] = {n: float((a - b).norm() / b.norm().clamp_min(1e-30)) for n, a, b in zip(names, keep["full"], keep["generic"])}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run(131072, 768)
    run(278528, 1024)
    run(16384, 768)
