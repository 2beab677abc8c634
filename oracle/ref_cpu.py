"""CPU baselines built from the REAL reference modules (``oracle/refimport.load_reference()``: /root/reference in the build
container, the verbatim ``oracle/_ref`` snapshot on the GPU box).  BASELINE INFRASTRUCTURE ONLY — bench.py's
``cpu_baseline`` / ``--impl reference`` legs are the only callers.

BASELINE.md §3: reference modules unmodified, ``device='cpu'``, reference dtypes, eager, ``RMSprop(lr=1e-4)``
(src/cli/train.py:41,50,189-190), ``torch.set_num_threads(os.cpu_count())``, seeded synthetic windows, warm-up then
``time.perf_counter`` over the timed steps.  The step is the reference loop body (train.py:246-284):
``optimizer.zero_grad(); outputs = model(inputs); loss = evaluator(...); loss.backward(); optimizer.step()``.

The diffusion denoiser does not exist in the reference (only trace: src/.gitignore:10).  Its CPU arm is therefore
COMPOSED here from reference parts — ``TransformerLayer`` x L at the benchmark width (TransformerBaseline.py:8-38, fp32),
``RegressionLossEvaluator.__call__`` as the loss, torch.optim.RMSprop — around the builder's stem / time MLP / head
(DESIGN.md D-1); labelled as such in every line that reports it.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import time
from typing import Callable, Dict

import torch
import torch.nn as nn

from . import ddpm as _ddpm
from . import models as _models
from .refimport import load_reference, reference_available

ARGS = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                          predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))


def available() -> bool:
    return reference_available()


def time_steps(step: Callable[[], None], warmup: int, steps: int) -> float:
    """Seconds per step."""
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class _Loop:
    """train.py:246-284 loop body around a reference model and the reference evaluator."""

    def __init__(self, ns, model: nn.Module, forward: Callable[[], Dict[str, torch.Tensor]], labels: Dict[str, torch.Tensor]):
        self.ns, self.model, self.forward, self.labels = ns, model, forward, labels
        self.opt = torch.optim.RMSprop(model.parameters(), lr=1e-4)
        self.ev = ns.RegressionLossEvaluator(dataset=None, split="train")
        self.n = 0

    def __call__(self):
        self.opt.zero_grad()
        loss = self.ev(None, dict(self.forward()), {k: v for k, v in self.labels.items()}, [], [], ARGS)
        loss.backward()
        self.opt.step()
        self.n += 1
        if self.n % 8 == 0:                 # the evaluator keeps every step's loss tensors until print_report (…py:188-221)
            self.ev = self.ns.RegressionLossEvaluator(dataset=None, split="train")
        return float(loss.detach())


def feedforward_loop(B: int, seed: int = 11):
    """BASELINE configs[0]: FeedForwardBaseline(23, 2, 50, all_frames, sigmoid, stride 5, hidden [512, 512])."""
    from .train import synthetic_batch
    ns = load_reference()
    torch.manual_seed(0)
    m = _quiet(ns.FeedForwardBaseline, 23, 2, 50, "all_frames", "sigmoid", 5, 10, hidden_dims=[512, 512])
    m.train()
    inputs, labels = synthetic_batch(B, 10, 23, 15, seed)
    return _Loop(ns, m, lambda: m(inputs), labels)


def groundlink_loop(B: int, T: int = 50, seed: int = 12):
    from .train import synthetic_batch
    ns = load_reference()
    torch.manual_seed(0)
    m = ns.Groundlink(23, 12, 10, "all_frames")
    m.train()
    inputs, labels = synthetic_batch(B, T, 23, 30, seed)
    return _Loop(ns, m, lambda: m(inputs), labels)


def transformer_forward(B: int, T: int = 200, seed: int = 13):
    """The TransformerBaseline layer stack + heads composed as its forward does (TransformerBaseline.py:104-148; the class's
    own forward needs key constants that do not exist), fp64, no_grad — the analyze pass."""
    ns = load_reference()
    torch.manual_seed(0)
    D = 23
    m = ns.TransformerBaseline(D, T)
    m.eval()
    g = torch.Generator().manual_seed(seed)
    x = {k: torch.randn(B, c, T, generator=g, dtype=torch.float64) for k, c in
         [("pos", D), ("vel", D), ("acc", D), ("comPos", 3), ("comVel", 3), ("comAcc", 3)]}

    def run():
        with torch.no_grad():
            vecs = torch.cat([x["pos"], x["vel"], x["acc"], x["comPos"], x["comVel"], x["comAcc"]], dim=1).transpose(1, 2)
            emb = m.temporal_embedding(torch.arange(vecs.size(1))).expand(B, T, m.temporal_embedding_dim)
            vecs = torch.cat([vecs, emb], dim=2)
            for layer in m.transformer_layers:
                vecs = layer(vecs)
            out = m.fc(vecs)
            blend = m.com_attention(vecs, vecs, x["comAcc"].transpose(1, 2))
            return m.contact_sigmoid(out[:, :, :2]), blend, out[:, :, 5:]
    return run


class ComposedDenoiser(nn.Module):
    """Builder's denoiser with the reference's own ``TransformerLayer`` as its layers."""

    def __init__(self, ns, c_in: int, frames: int, d: int, heads: int, ff: int, layers: int):
        super().__init__()
        self.d = d
        self.in_proj = nn.Linear(30 + c_in, d)
        self.time_mlp = nn.Sequential(nn.Linear(d, d), nn.SiLU(), nn.Linear(d, d))
        self.pos_embedding = nn.Parameter(torch.randn(frames, d) * 0.02)
        self.layers = nn.ModuleList([ns.TransformerLayer(d, heads, ff, 0.0, dtype=torch.float32) for _ in range(layers)])
        self.out_proj = nn.Linear(d, 30)

    def forward(self, cond, x_t, t):
        h = self.in_proj(torch.cat([x_t, cond], dim=-1))
        e = self.time_mlp(_models.sinusoidal_embedding(t, self.d))
        h = h + e.unsqueeze(1) + self.pos_embedding[: h.shape[1]].unsqueeze(0)
        for layer in self.layers:
            h = layer(h)
        return self.out_proj(h)


def denoiser_loop(B: int, c_in: int, frames: int, d: int, heads: int, ff: int, layers: int, seed: int = 1234):
    """One denoiser TRAINING step per call: random timesteps, q_sample, forward, reference loss, backward, RMSprop."""
    from .train import synthetic_batch
    ns = load_reference()
    torch.manual_seed(0)
    m = ComposedDenoiser(ns, c_in, frames, d, heads, ff, layers)
    m.train()
    sched = _ddpm.make_schedule()
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(B, frames, c_in, generator=g)
    _, labels = synthetic_batch(B, frames, 23, 30, seed + 1)
    x0 = torch.cat([labels[k] for k in (_models.COP, _models.FORCE, _models.TORQUE, _models.WRENCH)], dim=-1)

    def forward():
        t = torch.randint(0, 1000, (B,), generator=g)
        eps = torch.randn(B, frames, 30, generator=g)
        x_t = _ddpm.q_sample(sched, x0, t, eps)
        return _models.split30(m(cond, x_t, t))
    return _Loop(ns, m, forward, labels)


def denoiser_sampler(B: int, c_in: int, frames: int, d: int, heads: int, ff: int, layers: int, seed: int = 1234):
    """One reverse-diffusion (denoise) step per call on B windows: composed denoiser forward + posterior update."""
    ns = load_reference()
    torch.manual_seed(0)
    m = ComposedDenoiser(ns, c_in, frames, d, heads, ff, layers)
    m.eval()
    sched = _ddpm.make_schedule()
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(B, frames, c_in, generator=g)
    state = {"x": torch.randn(B, frames, 30, generator=g), "t": 999}

    def step():
        t = state["t"]
        with torch.no_grad():
            x0_hat = m(cond, state["x"], torch.full((B,), t))
            z = torch.randn(B * frames, 30, generator=g)
            state["x"] = _ddpm.posterior_step(sched, x0_hat.reshape(-1, 30), state["x"].reshape(-1, 30), t, z).view(B, frames, 30)
        state["t"] = t - 1 if t > 0 else 999
    return step


def host() -> dict:
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"cores": os.cpu_count() or 1, "cpu_model": model}
