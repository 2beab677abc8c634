"""Diagnostic (not a pytest): run one EncoderLayerPlan forward/backward and compare every
intermediate with a torch fp32 recomputation on the GPU.  python tools/diag_layer.py [d heads ff B T]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from inferbiomechanics_b200.engine import EncoderLayerPlan, _Buffers  # noqa: E402
from inferbiomechanics_b200.models.DiffusionDenoiser import _TransformerLayerParams  # noqa: E402
from inferbiomechanics_b200.params import ParamArena  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def main(d=128, heads=2, ff=256, B=3, T=50):
    torch.manual_seed(0)
    mod = _TransformerLayerParams(d, heads, ff).cuda()
    arena = ParamArena(list(mod.named_parameters()), torch.device("cuda"))
    plan = EncoderLayerPlan(arena, "", d, heads, ff)
    buf = _Buffers(torch.device("cuda"))
    st = buf.get((B,))
    M = B * T
    a = plan.alloc(buf, st, "L0", M, True)
    x = torch.randn(M, d, device="cuda").to(torch.bfloat16)
    dy = torch.randn(M, d, device="cuda").to(torch.bfloat16)
    y = plan.forward(x, a, M, B, T)
    sc = {k: torch.empty(M, w, dtype=torch.bfloat16, device="cuda") for k, w in
          (("ds", d), ("dh", ff), ("dx1", d), ("do", d), ("dqkv", 3 * d))}
    dx = torch.empty(M, d, dtype=torch.bfloat16, device="cuda")
    arena.zero_grad()
    # keep copies of scratch after each stage by re-running pieces is complex; instead recompute refs
    plan.backward(x, a, dy, sc, M, B, T, dx)
    torch.cuda.synchronize()

    # torch fp32 reference with the SAME bf16-rounded weights
    W = {n: arena.shadow_of(n).float().view(p.shape).clone().requires_grad_(True) if p.dim() == 2 else
         p.detach().clone().float().requires_grad_(True) for n, p in mod.named_parameters()}
    xr = x.float().clone().requires_grad_(True)
    qkv = xr @ W["multihead_attention.in_proj_weight"].t() + W["multihead_attention.in_proj_bias"]
    hd = d // heads
    q, k, v = [t.view(B, T, heads, hd).transpose(1, 2) for t in qkv.split(d, dim=-1)]
    p = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(hd), dim=-1)
    o = (p @ v).transpose(1, 2).reshape(M, d)
    s1 = o @ W["multihead_attention.out_proj.weight"].t() + W["multihead_attention.out_proj.bias"] + xr
    x1 = torch.nn.functional.layer_norm(s1, (d,), W["norm1.weight"], W["norm1.bias"])
    h = torch.relu(x1 @ W["feedforward.0.weight"].t() + W["feedforward.0.bias"])
    s2 = h @ W["feedforward.2.weight"].t() + W["feedforward.2.bias"] + x1
    x2 = torch.nn.functional.layer_norm(s2, (d,), W["norm2.weight"], W["norm2.bias"])
    for t in (qkv, o, s1, x1, h, s2):
        t.retain_grad()
    x2.backward(dy.float())
    print(f"d={d} heads={heads} ff={ff} B={B} T={T}")
    for name, got, ref in (("qkv", a["qkv"], qkv), ("o", a["o"], o), ("s1", a["s1"], s1), ("x1", a["x1"], x1), ("h", a["h"], h),
                           ("s2", a["s2"], s2), ("x2", y, x2)):
        print(f"  fwd {name:5s} rel err {rel(got, ref.detach()):.4f}")
    # backward scratch: ds holds ds1 at the end; dh, dx1, do, dqkv hold their final values
    for name, got, ref in (("dh(pre-relu)", sc["dh"], h.grad * (h > 0)), ("dx1", sc["dx1"], x1.grad), ("ds1", sc["ds"], s1.grad),
                           ("do", sc["do"], o.grad), ("dqkv", sc["dqkv"], qkv.grad), ("dx", dx, xr.grad)):
        print(f"  bwd {name:12s} rel err {rel(got, ref):.4f}")
    for n, prm in mod.named_parameters():
        print(f"  grad {n:45s} rel err {rel(prm.grad, W[n].grad):.4f}")


if __name__ == "__main__":
    main(*[int(v) for v in sys.argv[1:]])
