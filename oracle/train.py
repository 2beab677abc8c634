"""CPU port of one reference-shaped training step.  TEST INFRASTRUCTURE ONLY.

Shape of the step follows ``/root/reference/src/cli/train.py:240-284``:
``optimizer.zero_grad(); outputs = model(inputs); loss = evaluator(...); loss.backward();
optimizer.step()`` with ``torch.optim.RMSprop(lr=1e-4)`` (train.py:189-190; torch is the
reference's own dependency so its optimizer *is* the reference arithmetic).

Used by bench.py's ``cpu_baseline`` / ``--impl reference`` legs ("kind": "port") and by the
parity tests for multi-step training trajectories.  The model forwards are the functional
restatements in ``oracle/models.py``; autograd supplies the backward exactly as in the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Sequence

import torch

from . import ddpm as _ddpm
from . import loss as _loss
from . import models as _models

ALL = dict(grf=list(range(6)), cop=list(range(6)), moment=list(range(6)), wrench=list(range(12)))


def _linear_init(out_f: int, in_f: int, gen: torch.Generator, bias: bool = True):
    """nn.Linear.reset_parameters: kaiming_uniform(a=sqrt(5)) ⇒ U(-1/sqrt(in), 1/sqrt(in))."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound if bias else None
    return w, b


def init_feedforward(input_size: int, hidden: Sequence[int], output_size: int, seed: int) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    dims = [input_size] + list(hidden) + [output_size]
    sd = {}
    for i, (h0, h1) in enumerate(zip(dims[:-1], dims[1:])):
        w, b = _linear_init(h1, h0, gen)
        sd[f"net.{2 * i}.weight"], sd[f"net.{2 * i}.bias"] = w, b
    return sd


def init_denoiser(c_in: int, frames: int, d: int, ff: int, layers: int, seed: int) -> Dict[str, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, o, i):
        sd[name + ".weight"], sd[name + ".bias"] = _linear_init(o, i, gen)

    lin("in_proj", d, 30 + c_in)
    lin("time_mlp.0", d, d)
    lin("time_mlp.2", d, d)
    sd["pos_embedding"] = torch.randn(frames, d, generator=gen) * 0.02
    for l in range(layers):
        p = f"layers.{l}."
        w, b = _linear_init(3 * d, d, gen)
        sd[p + "multihead_attention.in_proj_weight"], sd[p + "multihead_attention.in_proj_bias"] = w, b
        lin(p + "multihead_attention.out_proj", d, d)
        lin(p + "feedforward.0", ff, d)
        lin(p + "feedforward.2", d, ff)
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = torch.ones(d)
            sd[p + n + ".bias"] = torch.zeros(d)
    lin("out_proj", 30, d)
    return sd


class PortTrainer:
    """Holds fp32 leaf parameters + torch.optim.RMSprop; ``step`` runs one training step."""

    def __init__(self, sd: Mapping[str, torch.Tensor], lr: float = 1e-4, opt: str = "rmsprop"):
        self.params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        cls = {"rmsprop": torch.optim.RMSprop, "adam": torch.optim.Adam, "sgd": torch.optim.SGD,
               "adagrad": torch.optim.Adagrad, "adadelta": torch.optim.Adadelta,
               "adamax": torch.optim.Adamax}[opt]
        self.opt = cls(list(self.params.values()), lr=lr)

    def _finish(self, out4: Dict[str, torch.Tensor], labels: Mapping[str, torch.Tensor], sel=ALL):
        res = _loss.regression_loss(out4, labels, sel["grf"], sel["cop"], sel["moment"], sel["wrench"])
        res["loss"].backward()
        self.opt.step()
        return res

    def step_feedforward(self, inputs, labels, activation: str, num_output_frames: int, sel=ALL):
        self.opt.zero_grad()
        out = _models.feedforward_forward(self.params, inputs, activation, num_output_frames)
        return self._finish(out, labels, sel)

    def step_groundlink(self, inputs, labels, sel=ALL):
        self.opt.zero_grad()
        out = _models.groundlink_forward(self.params, inputs)
        return self._finish(out, labels, sel)

    def step_denoiser(self, sched, cond, x0, t, eps, labels, num_layers: int, num_heads: int, sel=ALL):
        self.opt.zero_grad()
        x_t = _ddpm.q_sample(sched, x0, t, eps)
        x0_hat = _models.denoiser_forward(self.params, cond, x_t, t, num_layers, num_heads)
        return self._finish(_models.split30(x0_hat), labels, sel)


def synthetic_batch(B: int, F: int, num_dofs: int, hist_cols: int, seed: int, label_frames: int = None):
    """SURVEY §8d synthetic inputs: kinematics N(0,1); label forces N(0,1)*10; others N(0,1)."""
    from .windows import input_widths
    g = torch.Generator().manual_seed(seed)
    inputs = {k: torch.randn(B, F, c, generator=g) for k, c in input_widths(num_dofs, hist_cols).items()}
    Fo = F if label_frames is None else label_frames
    labels = {
        _loss.COP: torch.randn(B, Fo, 6, generator=g),
        _loss.FORCE: torch.randn(B, Fo, 6, generator=g) * 10.0,
        _loss.TORQUE: torch.randn(B, Fo, 6, generator=g),
        _loss.WRENCH: torch.randn(B, Fo, 12, generator=g),
    }
    return inputs, labels
