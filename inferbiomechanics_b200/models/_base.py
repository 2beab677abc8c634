"""Shared plumbing of the drop-in model classes: arena management and dict-input packing."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch
import torch.nn as nn

from .. import _lib, ops
from ..keys import MODEL_INPUT_ORDER
from ..params import ParamArena


class EngineModule(nn.Module):
    """nn.Module whose parameters are views into a flat HBM arena read by libibm_b200 kernels."""

    def _init_engine_state(self):
        self._arena: Optional[ParamArena] = None
        self._engine = None
        self._stage: Dict[tuple, torch.Tensor] = {}

    def _device(self) -> torch.device:
        p = next(self.parameters())
        if not p.is_cuda:
            raise _lib.IbmError(
                f"{type(self).__name__} runs only on a B200: move it to CUDA with .to('cuda') "
                "(inferbiomechanics_b200 has no CPU fallback)")
        return p.device

    def _build_engine(self, arena: ParamArena):
        raise NotImplementedError

    def engine(self):
        dev = self._device()
        if self._arena is None or not self._arena.intact():
            self._arena = ParamArena(list(self.named_parameters()), dev)
            self._engine = self._build_engine(self._arena)
            # module loop (the reference's autograd + DDP idiom): every data-parallel rank draws its own dropout masks, as
            # per-process torch RNGs do; the native Trainer re-seeds with its own seed and step count (Trainer.seed_rng)
            from ..parallel import world
            for attr in ("dropout_seed", "cnn_seed", "fc_seed"):
                if hasattr(self._engine, attr):
                    setattr(self._engine, attr, getattr(self._engine, attr) + world()[0])
        else:
            self._arena.sync_shadow()
        return self._engine

    @property
    def arena(self) -> ParamArena:
        self.engine()
        return self._arena

    # ---- dict → packed bf16 rows ---------------------------------------------------------------
    def _pack_dict(self, input: Dict[str, torch.Tensor], dst_bf16: torch.Tensor, F: int, frame_stride: int, win_extra: int,
                   col0: int, keys: Sequence[str] = MODEL_INPUT_ORDER) -> None:
        """torch.concat([...10 keys...], -1) of the reference (FeedForward…py:97-108), as one kernel
        writing bf16 rows.  CPU tensors are concatenated into one pinned staging buffer and moved with a
        single H2D copy (the reference issues one copy per forward too, after its concat)."""
        dev = dst_bf16.device
        ts = [input[k] for k in keys]
        B = ts[0].shape[0]
        n_rows = B * F
        if all(not t.is_cuda for t in ts):
            C = sum(t.shape[-1] for t in ts)
            key = ("pin", n_rows, C)
            if key not in self._stage:
                self._stage[key] = torch.empty(n_rows, C, dtype=torch.float32).pin_memory()
                self._stage[("dev",) + key[1:]] = torch.empty(n_rows, C, dtype=torch.float32, device=dev)
            pin, devbuf = self._stage[key], self._stage[("dev",) + key[1:]]
            torch.cat([t.reshape(n_rows, -1).to(torch.float32) for t in ts], dim=-1, out=pin)
            devbuf.copy_(pin, non_blocking=True)
            srcs = [devbuf]
        else:
            srcs = [t.to(dev, torch.float32, non_blocking=True).contiguous().view(n_rows, -1) for t in ts]
        ops.pack_inputs(srcs, n_rows, F, out_bf16=dst_bf16, frame_stride=frame_stride, win_extra=win_extra, col0=col0)
