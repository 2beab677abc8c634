"""One small invocation of the hot path on cuda:0, checked against the CPU oracle
(used by __graft_entry__.smoke(); the oracle is only the checker here)."""
from __future__ import annotations

import argparse

import torch


def run_smoke() -> None:
    from oracle import loss as ol
    from oracle import models as om
    from oracle.gen_golden import seeded_inputs, seeded_out_labels
    from oracle.seeded import seeded_state_dict

    from .data.window_store import WindowStore
    from .loss.RegressionLossEvaluator import RegressionLossEvaluator
    from .models.DiffusionDenoiser import DiffusionDenoiser
    from .models.FeedForwardRegressionBaseline import FeedForwardBaseline
    from .trainer import Trainer

    torch.cuda.set_device(0)
    args = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                              predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))
    # BASELINE config 1 shape: T=50, stride 5, 23 DOF, batch 32, hidden [512, 512], sigmoid
    D, T, s, B = 23, 50, 5, 32
    model = FeedForwardBaseline(D, 2, T, "all_frames", "sigmoid", s, 10, hidden_dims=[512, 512])
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, 1)
    model.load_state_dict(sd)
    model = model.cuda()
    inputs = seeded_inputs(B, T // s, D, s * 3, 2)
    _, labels = seeded_out_labels(B, T // s, 3)
    out = model(inputs)
    ev = RegressionLossEvaluator(None, "train", device="cuda")
    loss = ev(inputs, out, {k: v.clone() for k, v in labels.items()}, [], [], args)
    for p in model.parameters():
        p.grad = None
    loss.backward()
    ref_out = om.feedforward_forward(sd, inputs, "sigmoid", T // s)
    ref = ol.regression_loss(ref_out, labels, range(6), range(6), range(6), range(12))
    rel = abs(loss.item() - ref["loss"].item()) / abs(ref["loss"].item())
    assert rel < 2e-2, f"FeedForward loss mismatch vs oracle: {loss.item()} vs {ref['loss'].item()}"
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())

    # one native denoiser training step (tcgen05 GEMMs, attention, LN, fused loss, fused optimizer)
    den = DiffusionDenoiser(frames=10, d_model=128, num_heads=2, dim_feedforward=256, num_layers=2).cuda()
    store = WindowStore.synthetic(1024, 10, 1, 177, "all_frames", seed=4, trial_len=200)
    tr = Trainer(den, opt_type="rmsprop", lr=1e-4)
    res = tr.train_step(store, store.shard(0, 1)[:64])
    torch.cuda.synchronize()
    assert torch.isfinite(res[0]), "denoiser step produced a non-finite loss"
    print(f"smoke ok: feedforward loss {loss.item():.6f} (oracle {ref['loss'].item():.6f}), denoiser loss {res[0].item():.4f}")
