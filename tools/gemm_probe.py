#!/usr/bin/env python
"""Time individual ibm_gemm_bf16 configurations with CUDA events (tuning aid; also the command ncu wraps).

    python tools/gemm_probe.py                 # the shapes of one denoiser training step
    python tools/gemm_probe.py ffn2_dgrad      # a single named case (what ncu is pointed at)
"""
import sys

import torch

sys.path.insert(0, ".")
from inferbiomechanics_b200 import ops  # noqa: E402

M = 204800
CASES = {
    # name: (M, N, K, a_mn, b_mn, bias, act, aux_mode, out_f32, accumulate)
    "qkv_fwd": (M, 1536, 512, 0, 0, 1, "none", 0, 0, 0),
    "ffn1_fwd": (M, 2048, 512, 0, 0, 1, "relu", 0, 0, 0),
    "ffn2_fwd_res": (M, 512, 2048, 0, 0, 1, "none", 1, 0, 0),
    "ffn2_fwd_plain": (M, 512, 2048, 0, 0, 1, "none", 0, 0, 0),       # the same product without the residual tile (4 operand stages)
    "outproj_fwd_res": (M, 512, 512, 0, 0, 1, "none", 1, 0, 0),
    "outproj_dgrad": (M, 512, 512, 0, 1, 0, "none", 0, 0, 0),
    "ffn2_dgrad": (M, 2048, 512, 0, 1, 0, "relu", 2, 0, 0),
    "ffn2_dgrad_cs": (M, 2048, 512, 0, 1, 0, "relu", 2, 0, 0),       # + fused bias-gradient column sums
    "ffn2_dgrad_mask_cs": (M, 2048, 512, 0, 1, 0, "none", 0, 0, 0),  # ReLU derivative from the sign bitmask + column sums
    "ffn1_fwd_mask": (M, 2048, 512, 0, 0, 1, "relu", 0, 0, 0),       # forward that also writes the sign bitmask
    "ffn1_dgrad_res": (M, 512, 2048, 0, 1, 0, "none", 1, 0, 0),
    "qkv_dgrad_res": (M, 512, 1536, 0, 1, 0, "none", 1, 0, 0),
    "ffn1_wgrad": (2048, 512, M, 1, 1, 0, "none", 0, 1, 1),
    "ffn2_wgrad": (512, 2048, M, 1, 1, 0, "none", 0, 1, 1),
    "qkv_wgrad": (1536, 512, M, 1, 1, 0, "none", 0, 1, 1),
    "outproj_wgrad": (512, 512, M, 1, 1, 0, "none", 0, 1, 1),
}


def run(name, iters=10):
    m, n, k, a_mn, b_mn, bias, act, aux_mode, f32, acc = CASES[name]
    dev = "cuda"
    A = torch.randn((k, m) if a_mn else (m, k), device=dev).to(torch.bfloat16)
    B = torch.randn((k, n) if b_mn else (n, k), device=dev).to(torch.bfloat16)
    out = torch.zeros(m, n, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
    bv = torch.randn(n, device=dev) if bias else None
    aux = torch.randn(m, n, device=dev).to(torch.bfloat16) if aux_mode else None
    cs = torch.zeros(n, device=dev) if name.endswith("_cs") else None
    mask = torch.zeros(m, n // 8, dtype=torch.uint8, device=dev) if "_mask" in name else None
    mm = 0 if mask is None else (1 if act == "relu" else 2)
    f = lambda: ops.gemm(A, B, out, m, n, k, a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bv, act=act, aux=aux, aux_mode=aux_mode,
                         accumulate=bool(acc), colsum=cs, mask=mask, mask_mode=mm)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    byts = (A.numel() + B.numel()) * 2 + out.numel() * out.element_size() + (aux.numel() * 2 if aux is not None else 0)
    # the vendor library on the bare product (no bias / activation / residual / mask / column sums): what torch.matmul (cuBLASLt)
    # needs for the same M x N x K with the same operand layouts — the library path would add elementwise kernels on top
    lib = ""
    if "--cublas" in sys.argv:
        Af = A.t() if a_mn else A
        Bf = B if b_mn else B.t()
        ref = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            torch.matmul(Af, Bf, out=ref)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(Af, Bf, out=ref)
        e1.record()
        torch.cuda.synchronize()
        lms = e0.elapsed_time(e1) / iters
        lib = f"   | cuBLAS bare bf16 product {lms * 1e3:8.1f} us {2.0 * m * n * k / lms / 1e9:8.1f} TFLOP/s"
    print(f"{name:18s} {m}x{n}x{k}  {ms * 1e3:8.1f} us  {2.0 * m * n * k / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:7.0f} GB/s (algorithmic){lib}")


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(CASES)
    for nm in names:
        run(nm, iters=10)
