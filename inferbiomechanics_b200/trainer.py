"""B200-native training step for the reference's training loop shape
(/root/reference/src/cli/train.py:240-284): zero_grad → forward → RegressionLossEvaluator → backward
→ (DDP allreduce) → optimizer.step, without autograd, without per-step host syncs:

    grads.zero_ (one memset of the flat arena)
    window packer  → bf16 activations            (1-2 kernels)
    forward engine → fp32 outputs                (tcgen05 GEMMs + fused elementwise)
    fused loss fwd → fp32[40] result on device   (1 kernel)
    fused loss bwd → bf16 d loss/d out           (1 kernel)
    backward engine → flat fp32 grad arena       (bucketed NCCL allreduce fired layer by layer)
    fused optimizer → params + state + bf16 shadow (1 kernel, 1/W folded in)

Works for ``FeedForwardBaseline`` and ``DiffusionDenoiser``.  The drop-in classes remain usable with
the reference's own autograd loop; this is the path bench.py times.
"""
from __future__ import annotations

import argparse
from typing import Dict, List, Optional

import torch

from . import ops, parallel
from .data.window_store import WindowStore
from .diffusion import GaussianDiffusion
from .keys import LOSS_QUANTITIES
from .loss.RegressionLossEvaluator import COP_FORCE_THRESHOLD, component_weights
from .models.DiffusionDenoiser import DiffusionDenoiser
from .models.FeedForwardRegressionBaseline import FeedForwardBaseline

ALL_COMPONENTS = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                                    predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))


def _views30(rows: torch.Tensor):
    """(B,F,>=30) rows30 tensor → the four quantity views (cop, force, torque, wrench)."""
    return [rows[:, :, 0:6], rows[:, :, 6:12], rows[:, :, 12:18], rows[:, :, 18:30]]


def _views_ff(x: torch.Tensor, B: int, Fo: int):
    """FeedForward output/grad layout (FeedForward…py:116-121): quantity-then-frame blocks of a [B, >=30*Fo] row."""
    def blk(a, b, c):
        return x[:, a * Fo:b * Fo].unflatten(1, (Fo, c))
    return [blk(0, 6, 6), blk(6, 12, 6), blk(12, 18, 6), blk(18, 30, 12)]


class Trainer:
    def __init__(self, model, opt_type: str = "rmsprop", lr: float = 1e-4, args: Optional[argparse.Namespace] = None,
                 diffusion: Optional[GaussianDiffusion] = None, bucket_mb: float = 8.0, seed: int = 0):
        self.model = model
        self.eng = model.engine()
        self.arena = model.arena
        self.opt_type, self.lr = opt_type, lr
        self.weights = component_weights(args or ALL_COMPONENTS)
        self.rank, self.world = parallel.world()
        self.step_count = 0
        self.seed = seed
        self.is_denoiser = isinstance(model, DiffusionDenoiser)
        if self.is_denoiser:
            self.diffusion = diffusion or GaussianDiffusion(device=self.arena.device)
        n = self.arena.total
        dev = self.arena.device
        self.state0 = torch.zeros(n, device=dev) if opt_type != "sgd" else None
        self.state1 = torch.zeros(n, device=dev) if opt_type in ("adam", "adadelta", "adamax") else None
        self.results: List[torch.Tensor] = []              # fp32[40] per step, device-resident
        self._result_ring = [torch.zeros(40, device=dev) for _ in range(8)]
        # layer groups for the bucketed allreduce
        bounds = self._group_boundaries()
        self.bucketer = parallel.GradBucketer(self.arena.grad, parallel.make_buckets(bounds, n, int(bucket_mb * (1 << 20) / 4)))
        if self.is_denoiser:
            self.eng.bucket_hook = lambda l: self.bucketer.group_done(l + 1)      # group 0 = stem, l+1 = layer l, L+1 = head
        else:
            self.eng.bucket_hook = lambda i: self.bucketer.group_done(i)          # group i = Linear layer i
        # DDP constructor semantics: everyone starts from rank 0's parameters (train.py:175)
        self.bucketer.broadcast_(self.arena.master)
        self.arena.sync_shadow(force=True)
        self._lab: Dict[int, torch.Tensor] = {}
        self._gen = torch.Generator(device=dev).manual_seed(seed + 7919 * self.rank)

    def _group_boundaries(self) -> List[int]:
        offs = self.arena.offsets
        if self.is_denoiser:
            first = [offs[f"layers.{l}.multihead_attention.in_proj_weight"][0] for l in range(self.model.num_layers)]
            return [0] + first + [offs["out_proj.weight"][0]]
        return [offs[w][0] for (w, _, _, _) in self.eng.layers]

    # ---- one optimisation step on windows idx of a store -------------------------------------------
    def train_step(self, store: WindowStore, idx: torch.Tensor) -> torch.Tensor:
        """Returns the device-resident fp32[40] result (loss at [0]); nothing is synchronised."""
        B = idx.numel()
        lab = self._labels(store, idx)
        self.arena.zero_grad()
        self.bucketer.begin_step()
        if self.is_denoiser:
            out = self._denoiser_forward(store, idx, lab, B)
            outs, labs = _views30(out.view(B, self.eng.F, 32)), _views30(lab)
            gviews = _views30(self.eng.dout(B).view(B, self.eng.F, 32))
        else:
            store.pack_feedforward(idx, self.eng.input_buffer(B))
            out = self.eng.forward(B)
            Fo = self.model.num_output_frames
            outs, labs = _views_ff(out, B, Fo), _views30(lab)
            gviews = _views_ff(self.eng.dout_buffer(B), B, Fo)
        result = self._result_ring[self.step_count % len(self._result_ring)]
        ops.regression_loss_fwd(outs, labs, self.weights, COP_FORCE_THRESHOLD, result=result)
        ops.regression_loss_bwd(outs, labs, self.weights, gviews, threshold=COP_FORCE_THRESHOLD)
        if self.is_denoiser:
            self.eng.backward(B)
        else:
            self._ff_backward(B)
        self.bucketer.finish()
        self.optimizer_step()
        return result

    def _labels(self, store: WindowStore, idx: torch.Tensor) -> torch.Tensor:
        B = idx.numel()
        if B not in self._lab:
            self._lab[B] = torch.empty(B, store.Fo, 30, dtype=torch.float32, device=self.arena.device)
        return store.labels(idx, self._lab[B])

    def _denoiser_forward(self, store, idx, lab, B):
        eng = self.eng
        xc = eng.xc(B, True)
        store.pack_rows(idx, xc, col0=30)
        t = eng.t_buffer(B, True)
        t.copy_(torch.randint(0, self.diffusion.T, (B,), device=t.device, generator=self._gen, dtype=torch.int32))
        # x_t = q_sample(x0 = labels rows30, t, eps ~ Philox) written straight into the concat buffer as bf16
        self.diffusion.q_sample(lab.view(B, -1), t, None, xt_bf16=xc, bf16_ld=eng.ld_in, seed=self.seed + self.rank,
                                offset=self.step_count)
        return eng.forward(B, train=True)

    def _ff_backward(self, B: int) -> None:
        eng = self.eng
        # FeedForwardEngine.backward walks the layers from last to first; fire buckets as groups complete
        eng.backward(B)
        # (the MLP is 3 GEMM triples: the allreduce is one or two buckets, flushed in finish())

    def optimizer_step(self) -> None:
        self.step_count += 1
        ops.optimizer_step(self.opt_type, self.arena.master, self.arena.grad, self.state0, self.state1, self.arena.shadow,
                           self.lr, 1.0 / self.world, self.step_count)
        self.arena.mark_shadow_fresh()

    # ---- evaluation (no_grad forward + loss), used by analyze / dev-eval ------------------------------
    @torch.no_grad()
    def eval_step(self, store: WindowStore, idx: torch.Tensor) -> torch.Tensor:
        B = idx.numel()
        lab = self._labels(store, idx)
        if self.is_denoiser:
            raise NotImplementedError("denoiser evaluation = reverse sampling; use GaussianDiffusion.sample")
        store.pack_feedforward(idx, self.eng.input_buffer(B))
        out = self.eng.forward(B)
        return ops.regression_loss_fwd(_views_ff(out, B, self.model.num_output_frames), _views30(lab), self.weights)
