#!/usr/bin/env python
"""Time the non-GEMM kernels of one denoiser training step in isolation (CUDA events, inputs far larger
than L2) and print achieved algorithmic GB/s against the measured HBM peak.  Also the command ncu wraps.

    python tools/kernel_probe.py                # all cases
    python tools/kernel_probe.py attn_bwd       # one case
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from inferbiomechanics_b200 import ops  # noqa: E402

B, F, D, H, FF = 4096, 50, 512, 8, 2048
M = B * F
dev = "cuda"


def peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


def case(name):
    if name == "attn_fwd":
        qkv, o = bf(M, 3 * D), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.attention_fwd_fused(qkv, D, o, B, F, H, D // H, 0.125)), 2 * M * 4 * D
    if name == "attn_bwd":
        qkv, do, dqkv = bf(M, 3 * D), bf(M, D), torch.empty(M, 3 * D, dtype=torch.bfloat16, device=dev)
        db = torch.zeros(3 * D, device=dev)
        return (lambda: ops.attention_bwd(qkv, D, do, dqkv, B, F, H, D // H, 0.125, dbias=db)), 2 * M * 7 * D
    if name == "ln_fwd":
        s, y = bf(M, D), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
        return (lambda: ops.layernorm_fwd(s, y, g, b, M, D, mean=mean, rstd=rstd)), 2 * M * 2 * D + 8 * M
    if name == "ln_bwd":
        s, dy, ds = bf(M, D), bf(M, D), torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        g = torch.ones(D, device=dev)
        mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
        dg, db, dc = (torch.zeros(D, device=dev) for _ in range(3))
        return (lambda: ops.layernorm_bwd(dy, s, g, mean, rstd, M, D, ds, dg, db, dc)), 2 * M * 3 * D + 8 * M
    if name == "colsum_ffn":
        x, out = bf(M, FF), torch.zeros(FF, device=dev)
        return (lambda: ops.colsum(x, M, FF, out)), 2 * M * FF
    if name in ("loss_fwd", "loss_bwd"):
        Bb = 16384                                   # 819 200 rows: out + labels = 203 MB > L2
        rows = torch.randn(Bb * F, 32, device=dev)
        lab = torch.randn(Bb, F, 30, device=dev) * 5
        v = lambda x: [x[..., 0:6], x[..., 6:12], x[..., 12:18], x[..., 18:30]]
        outs, labs = v(rows.view(Bb, F, 32)), v(lab)
        w = [1.0] * 30
        if name == "loss_fwd":
            res = torch.zeros(40, device=dev)
            return (lambda: ops.regression_loss_fwd(outs, labs, w, result=res)), Bb * F * 240
        g16 = v(torch.zeros(Bb, F, 32, dtype=torch.bfloat16, device=dev))
        return (lambda: ops.regression_loss_bwd(outs, labs, w, g16)), Bb * F * 300
    # ---- TransformerBaseline analysis shapes (BASELINE configs[4]): 2048 windows x T=200, d=108 -> 112, 3 heads 36 -> 48
    if name == "attn_fwd_t200":
        Bt, Tt, Ht, hp = 2048, 200, 3, 48
        qkv, o = bf(Bt * Tt, 3 * Ht * hp), torch.empty(Bt * Tt, Ht * hp, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.attention_fwd_fused(qkv, Ht * hp, o, Bt, Tt, Ht, hp, 1 / 6.0)), 2 * Bt * Tt * 4 * Ht * hp
    if name == "attn_com_blend_t200":
        Bt, Tt = 2048, 200
        q, k, v, o = bf(Bt * Tt, 112), bf(Bt * Tt, 112), bf(Bt * Tt, 8), torch.empty(Bt * Tt, 8, dtype=torch.bfloat16, device=dev)
        return (lambda: ops.attention_fwd(q, k, v, o, Bt, Tt, 1, 112, 8, 1.0)), 2 * Bt * Tt * (112 + 112 + 8 + 8)
    if name == "ln_fwd_d108":
        Mt = 2048 * 200
        s_, y = bf(Mt, 112), torch.empty(Mt, 112, dtype=torch.bfloat16, device=dev)
        g, b = torch.ones(108, device=dev), torch.zeros(108, device=dev)
        return (lambda: ops.layernorm_fwd(s_, y, g, b, Mt, 108)), 2 * Mt * 2 * 112
    raise KeyError(name)


CASES = ["attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd", "colsum_ffn", "loss_fwd", "loss_bwd", "attn_fwd_t200", "attn_com_blend_t200",
         "ln_fwd_d108"]


def run(name, iters=10):
    """`iters` launches captured in a CUDA graph and replayed: host-side ctypes marshalling (50-100 us per call, longer
    than some of these kernels) stays out of the number."""
    f, byts = case(name)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            f()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gbs = byts / ms / 1e6
    print(f"{name:12s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s algorithmic  ({100 * gbs / peak():.1f}% of measured HBM peak)")


if __name__ == "__main__":
    names = sys.argv[1:] or CASES
    for nm in names:
        run(nm, iters=3 if len(sys.argv) > 1 else 10)
