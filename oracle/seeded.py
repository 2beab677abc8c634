"""Deterministic seeded weights / inputs shared by gen_golden.py and the tests.
TEST INFRASTRUCTURE ONLY.

Golden fixtures must stay small, so instead of storing multi-megabyte state dicts they store the
seed; both the generator script (which loads these weights into the imported reference modules)
and the tests (which load them into the CUDA-backed modules / the oracle restatement) rebuild
the same tensors with this one procedure.  torch's CPU generator is deterministic for a given
torch version; the GPU box runs the same image.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence

import torch


def seeded_state_dict(shapes: Mapping[str, Sequence[int]], seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shp = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.long)
            continue
        x = torch.randn(shp, generator=g, dtype=torch.float64)
        if name.endswith("running_var"):
            x = x.abs() + 0.5
        elif "norm" in name and name.endswith("weight"):
            x = 1.0 + 0.1 * x
        elif name.endswith("bias") or name.endswith("in_proj_bias") or name.endswith("running_mean"):
            x = 0.1 * x
        elif len(shp) >= 2:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            x = x / math.sqrt(fan_in)
        else:
            x = 0.1 * x
        sd[name] = x.to(dtype)
    return sd


def seeded_tensor(shape, seed: int, scale: float = 1.0, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(tuple(shape), generator=g, dtype=torch.float64) * scale).to(dtype)


def strided_sample(x: torch.Tensor, step: int = 97) -> torch.Tensor:
    return x.detach().reshape(-1)[::step].clone()
