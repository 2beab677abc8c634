"""Functional CPU restatement of the reference model forwards.  TEST INFRASTRUCTURE ONLY.

Each function takes a ``state_dict``-shaped mapping (the reference's own key names)
so that weights from the imported reference modules can be fed straight in, and re-states the
math with explicit matmuls (no nn.Linear / nn.Conv1d / nn.MultiheadAttention objects):

* ``feedforward_forward``  ← ``src/models/FeedForwardRegressionBaseline.py:80-121`` (+ layer
                             construction 65-77: [Dropout][BatchNorm1d] Linear act, no act last)
* ``groundlink_forward``   ← ``src/models/Groundlink.py:105-156`` (Conv1d k=7 replicate pad + ELU
                             ×4, per-frame MLP 256→256→256→30, 41-62)
* ``transformer_layer``    ← ``src/models/TransformerBaseline.py:8-38`` (post-LN encoder layer)
* ``transformer_forward``  ← ``src/models/TransformerBaseline.py:104-148`` (inputs are (B,C,T))
* ``denoiser_forward``     ← builder-owned DDPM denoiser (NOT in the reference — see ddpm.py)

Pinned by tests/golden/{ff,groundlink,transformer}_*.npz generated from the imported reference.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional

import torch

from .windows import INPUT_ORDER

COP = "groundContactCenterOfPressureInRootFrame"
FORCE = "groundContactForceInRootFrame"
TORQUE = "groundContactTorqueInRootFrame"
WRENCH = "groundContactWrenchesInRootFrame"

_ACT = {
    "relu": lambda x: torch.clamp_min(x, 0.0),
    "tanh": torch.tanh,
    "sigmoid": lambda x: 1.0 / (1.0 + torch.exp(-x)),
}


def concat_inputs(inputs: Mapping[str, torch.Tensor]) -> torch.Tensor:
    return torch.cat([inputs[k] for k in INPUT_ORDER], dim=-1)


def split30(x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """(B,F,30) → 4 keys; Groundlink.py:151-156 channel order."""
    return {COP: x[..., 0:6], FORCE: x[..., 6:12], TORQUE: x[..., 12:18], WRENCH: x[..., 18:30]}


def feedforward_forward(sd: Mapping[str, torch.Tensor], inputs: Mapping[str, torch.Tensor],
                        activation: str, num_output_frames: int,
                        batchnorm: bool = False, dropout: bool = False, training: bool = False,
                        new_stats: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
    """``training`` with ``batchnorm``: nn.BatchNorm1d's training mode (FeedForward…py:71-72) — normalise with the batch
    mean / biased batch variance; the running statistics would become (1-0.1)·running + 0.1·(mean, UNBIASED variance),
    returned through ``new_stats`` (keyed like the state_dict) instead of being written into ``sd``."""
    x = concat_inputs(inputs)
    B = x.shape[0]
    x = x.reshape(B, -1)
    # nn.Sequential positions: per layer [Dropout?][BatchNorm1d?] Linear [act]
    lin_keys = sorted({int(k.split(".")[1]) for k in sd if k.endswith(".weight") and sd[k].dim() == 2})
    for li, pos in enumerate(lin_keys):
        if batchnorm:  # BN on the layer INPUT (FeedForward…py:71-72)
            p = pos - 1
            rm, rv = sd[f"net.{p}.running_mean"], sd[f"net.{p}.running_var"]
            if training:
                mean, var = x.mean(0), x.var(0, unbiased=False)
                if new_stats is not None:
                    n = x.shape[0]
                    new_stats[f"net.{p}.running_mean"] = (0.9 * rm + 0.1 * mean).detach()
                    new_stats[f"net.{p}.running_var"] = (0.9 * rv + 0.1 * var * n / (n - 1)).detach()
                x = (x - mean) / torch.sqrt(var + 1e-5) * sd[f"net.{p}.weight"] + sd[f"net.{p}.bias"]
            else:
                x = (x - rm) / torch.sqrt(rv + 1e-5) * sd[f"net.{p}.weight"] + sd[f"net.{p}.bias"]
        x = x @ sd[f"net.{pos}.weight"].t() + sd[f"net.{pos}.bias"]
        if li < len(lin_keys) - 1:
            x = _ACT[activation](x)
    Fo = num_output_frames
    return {
        COP: x[:, 0 * Fo:6 * Fo].reshape(B, Fo, 6),
        FORCE: x[:, 6 * Fo:12 * Fo].reshape(B, Fo, 6),
        TORQUE: x[:, 12 * Fo:18 * Fo].reshape(B, Fo, 6),
        WRENCH: x[:, 18 * Fo:30 * Fo].reshape(B, Fo, 12),
    }


def _elu(x: torch.Tensor) -> torch.Tensor:
    return torch.where(x > 0, x, torch.expm1(x))


def conv1d_k_replicate(x_btc: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """x (B,T,Cin), w (Cout,Cin,K) → (B,T,Cout); replicate padding K//2 (Groundlink.py:41)."""
    B, T, Cin = x_btc.shape
    K = w.shape[-1]
    idx = (torch.arange(T).unsqueeze(1) + torch.arange(K).unsqueeze(0) - K // 2).clamp(0, T - 1)  # (T,K)
    cols = x_btc[:, idx, :]                              # (B,T,K,Cin)
    y = torch.einsum("btkc,ock->bto", cols, w)
    return y + b if b is not None else y


def groundlink_forward(sd: Mapping[str, torch.Tensor], inputs: Mapping[str, torch.Tensor],
                       output_data_format: str = "all_frames") -> Dict[str, torch.Tensor]:
    x = concat_inputs(inputs)                            # (B,T,C); Flatten(2,-1) is a no-op on 3-D
    for i in (1, 4, 7, 10):                              # cnn.{1,4,7,10}; dropout p=0
        x = _elu(conv1d_k_replicate(x, sd[f"cnn.{i}.weight"], sd[f"cnn.{i}.bias"]))
    if output_data_format != "all_frames":
        x = x[:, -1:, :]                                 # Groundlink.py:147-148 (after the CNN)
    for i in (2, 5):                                     # eval mode: Dropout(0.2) inactive
        x = _elu(x @ sd[f"fc.{i}.weight"].t() + sd[f"fc.{i}.bias"])
    x = x @ sd["fc.8.weight"].t()
    return split30(x)


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _identity(x: torch.Tensor) -> torch.Tensor:
    return x


def bf16_ste(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 in the forward value, identity in the gradient (straight-through).  Passed as ``rnd`` to the
    layer functions below it re-creates, in fp32 arithmetic, the points where the CUDA path STORES an activation as
    bf16 (after every GEMM epilogue, softmax probabilities, LayerNorm outputs).  With bf16-rounded weights on top,
    the oracle then takes the same ReLU gates as the kernels, so gradients can be compared at a tight bar instead of
    the loose one a pure-fp32 oracle forces (a flipped gate changes its gradient entries completely)."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


class _RoundBoth(torch.autograd.Function):
    """bf16 rounding of the value in forward AND of the incoming gradient in backward."""

    @staticmethod
    def forward(ctx, x, fwd):
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype), None


def bf16_both(x: torch.Tensor) -> torch.Tensor:
    """``bf16_ste`` plus the mirror image in backward: the CUDA path also STORES the gradient of every such activation
    as bf16 (dqkv, do, ds, dh, dx1, dx), so the gradient flowing back through the storage point is rounded too."""
    return _RoundBoth.apply(x, True)


def bf16_grad(x: torch.Tensor) -> torch.Tensor:
    """Identity forward, bf16-rounded gradient (the fp32 x0_hat output whose loss gradient the path casts to bf16)."""
    return _RoundBoth.apply(x, False)


def bf16_weights(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """GEMM weight matrices rounded to bf16 (the kernels' shadow arena); biases, LayerNorm parameters and the
    positional table stay fp32 (read from the master arena).  Straight-through, so gradients flow to ``sd``."""
    return {k: (bf16_ste(v) if v.dim() == 2 and k != "pos_embedding" else v) for k, v in sd.items()}


def multihead_attention(x: torch.Tensor, in_w, in_b, out_w, out_b, num_heads: int, rnd=_identity) -> torch.Tensor:
    B, T, d = x.shape
    hd = d // num_heads
    qkv = rnd(x @ in_w.t() + in_b)
    q, k, v = qkv.split(d, dim=-1)
    sh = lambda t: t.reshape(B, T, num_heads, hd).transpose(1, 2)      # (B,H,T,hd)
    q, k, v = sh(q), sh(k), sh(v)
    s = (q @ k.transpose(-2, -1)) / math.sqrt(hd)
    p = rnd(torch.softmax(s, dim=-1))
    o = rnd((p @ v).transpose(1, 2).reshape(B, T, d))
    return o @ out_w.t() + out_b


def transformer_layer(sd: Mapping[str, torch.Tensor], prefix: str, x: torch.Tensor, num_heads: int, rnd=_identity,
                      gate: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``gate`` (0/1 tensor shaped like the FFN hidden activation): replaces the ReLU's own decision — the test harness
    passes the gates the CUDA path took, so that a gradient comparison is not dominated by the few near-zero
    pre-activations whose sign differs between two roundings of the same forward."""
    g = lambda n: sd[prefix + n]
    a = multihead_attention(x, g("multihead_attention.in_proj_weight"), g("multihead_attention.in_proj_bias"),
                            g("multihead_attention.out_proj.weight"), g("multihead_attention.out_proj.bias"),
                            num_heads, rnd)
    x = rnd(layer_norm(rnd(x + a), g("norm1.weight"), g("norm1.bias")))
    pre = x @ g("feedforward.0.weight").t() + g("feedforward.0.bias")
    h = rnd(torch.clamp_min(pre, 0.0) if gate is None else pre * gate)
    f = h @ g("feedforward.2.weight").t() + g("feedforward.2.bias")
    return rnd(layer_norm(rnd(x + f), g("norm2.weight"), g("norm2.bias")))


def transformer_forward(sd: Mapping[str, torch.Tensor], x: Mapping[str, torch.Tensor], num_layers: int,
                        num_heads: int) -> Dict[str, torch.Tensor]:
    """TransformerBaseline.forward (104-148).  Inputs are (B,C,T); needs keys comPos/comVel/comAcc
    that the reference's own key class lacks — supplied by the harness."""
    vecs = torch.cat([x["pos"], x["vel"], x["acc"], x["comPos"], x["comVel"], x["comAcc"]], dim=1).transpose(1, 2)
    B, T, _ = vecs.shape
    emb = sd["temporal_embedding.embedding.weight"][:T].unsqueeze(0).expand(B, T, -1)
    vecs = torch.cat([vecs, emb], dim=2)
    for l in range(num_layers):
        vecs = transformer_layer(sd, f"transformer_layers.{l}.", vecs, num_heads)
    out = vecs @ sd["fc.weight"].t() + sd["fc.bias"]
    q = vecs @ sd["com_attention.query_linear.weight"].t() + sd["com_attention.query_linear.bias"]
    k = vecs @ sd["com_attention.key_linear.weight"].t() + sd["com_attention.key_linear.bias"]
    w = torch.softmax(q @ k.transpose(-2, -1), dim=-1)             # no 1/sqrt(d) (…:59-70)
    blend = w @ x["comAcc"].transpose(1, 2)
    return {
        "contact": (1.0 / (1.0 + torch.exp(-out[:, :, :2]))).transpose(1, 2),
        "comAcc": blend.transpose(1, 2),
        "contactForces": out[:, :, 5:].transpose(1, 2),
    }


# ---------------------------------------------------------------------------------------------
# Builder-owned denoiser (spec frozen in DESIGN.md §D-1).  NOT from the reference.
# ---------------------------------------------------------------------------------------------

def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32)
    omega = torch.exp(-math.log(10000.0) * k / half)
    a = t.to(torch.float32).unsqueeze(1) * omega.unsqueeze(0)
    return torch.cat([torch.sin(a), torch.cos(a)], dim=1)


def denoiser_forward(sd: Mapping[str, torch.Tensor], cond: torch.Tensor, x_t: torch.Tensor, t: torch.Tensor,
                     num_layers: int, num_heads: int, rnd=_identity, gates=None) -> torch.Tensor:
    """cond (B,F,C_in) packed kinematics; x_t (B,F,30); t (B,) int64 → x0_hat (B,F,30).
    ``rnd`` = ``bf16_ste`` / ``bf16_both`` (with ``sd = bf16_weights(sd)``) mirrors the kernels' bf16 storage points;
    ``gates``: one 0/1 tensor (B,F,ff) per layer, see ``transformer_layer``."""
    d = sd["in_proj.weight"].shape[0]
    h = rnd(rnd(torch.cat([x_t, cond], dim=-1)) @ sd["in_proj.weight"].t() + sd["in_proj.bias"])
    e = rnd(sinusoidal_embedding(t, d))
    e = rnd(e @ sd["time_mlp.0.weight"].t() + sd["time_mlp.0.bias"])
    e = rnd(e * (1.0 / (1.0 + torch.exp(-e))))                              # SiLU
    e = rnd(e @ sd["time_mlp.2.weight"].t() + sd["time_mlp.2.bias"])
    h = rnd(h + e.unsqueeze(1) + sd["pos_embedding"][: h.shape[1]].unsqueeze(0))
    for l in range(num_layers):
        h = transformer_layer(sd, f"layers.{l}.", h, num_heads, rnd, None if gates is None else gates[l])
    out = h @ sd["out_proj.weight"].t() + sd["out_proj.bias"]
    return bf16_grad(out) if rnd is bf16_both else out
