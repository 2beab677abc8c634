"""Auxiliary measurements of bench.py for the BASELINE configs other than the headline one (which is the
denoiser training step): configs[0] FeedForward training (reference shape B=32 and a GPU-sized batch),
Groundlink training, and configs[4] the TransformerBaseline analysis pass over a long-window (T=200) stream.

Every leg is timed with CUDA events on the launching stream after warm-up, through the drop-in classes /
Trainer (the public calls), and reports its algorithmic FLOPs (SURVEY §8d figures) next to the time.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _per_kernel(fn) -> Dict[str, dict]:
    """One extra instrumented pass: CUDA events around every C-ABI launch (live, no profiler)."""
    real_call = ops.call
    recs = []

    def timed_call(name, *a):
        e0, e1 = _events()
        e0.record()
        real_call(name, *a)
        e1.record()
        recs.append((name, e0, e1))

    ops.call = timed_call
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        ops.call = real_call
    per: Dict[str, list] = {}
    for name, e0, e1 in recs:
        p = per.setdefault(name, [0, 0.0])
        p[0] += 1
        p[1] += e0.elapsed_time(e1)
    return {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])}


def feedforward_train_leg(dev, world: int, steps: int = 20) -> dict:
    """BASELINE configs[0] shape on the GPU: FeedForwardBaseline 1470->512->512->300 (D=23, T=50, stride 5, sigmoid,
    RMSprop 1e-4; FeedForward...py:20-78, train.py:183-197), one Trainer.train_step per step (packer -> 3 tcgen05 GEMMs
    -> fused loss -> backward -> fused optimizer).  5 505 024 FLOP/window fwd+bwd."""
    from .data.window_store import WindowStore
    from .models.FeedForwardRegressionBaseline import FeedForwardBaseline
    from .trainer import Trainer
    torch.manual_seed(0)
    model = FeedForwardBaseline(23, 2, 50, "all_frames", "sigmoid", 5, 10, hidden_dims=[512, 512]).to(dev)
    trainer = Trainer(model, opt_type="rmsprop", lr=1e-4)
    out = {}
    for B in (32, 16384):
        store = WindowStore.synthetic(B * 2, 50, 5, 147, "all_frames", seed=77, device=dev)
        idx = store.shard(0, 1)
        batches = [idx[:B], idx[B:2 * B]]
        for i in range(3):
            trainer.train_step(store, batches[i % 2])
        e0, e1 = _events()
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            trainer.train_step(store, batches[i % 2])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[f"batch_{B}"] = {"windows_per_s": world * B / (ms * 1e-3), "ms_per_step": ms,
                             "tflops": 5505024.0 * B / (ms * 1e-3) / 1e12}
        del store
    out["note"] = "configs[0] model on the GPU; batch 32 is launch-latency bound (about 20 launches per step), batch 16384 is the throughput shape"
    return out


def groundlink_train_leg(dev, world: int, steps: int = 10, B: int = 4096, T: int = 50) -> dict:
    """Groundlink (Groundlink.py:20-156) training: implicit-GEMM temporal CNN + per-frame MLP, forward + backward with
    Dropout(0.2) active, 110.0 MFLOP/window forward at T=50, ~3x for the step.  Two paths over the same
    kernels: the native ``Trainer`` step (window store -> padded-row packer -> fused loss -> fused RMSprop) and the reference's
    loop shape (module forward -> evaluator -> loss.backward() -> torch.optim.RMSprop.step())."""
    import argparse
    from .data.window_store import WindowStore
    from .keys import LOSS_QUANTITIES, MODEL_INPUT_ORDER
    from .loss.RegressionLossEvaluator import RegressionLossEvaluator
    from .models.Groundlink import Groundlink
    from .trainer import Trainer
    torch.manual_seed(0)
    out = {"batch": B, "frames": T}

    def timed(step):
        for _ in range(3):
            step()
        e0, e1 = _events()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"windows_per_s": world * B / (ms * 1e-3), "ms_per_step": ms, "tflops_model": 3 * 110.016e6 * B / (ms * 1e-3) / 1e12}

    store = WindowStore.synthetic(2 * B, T, 1, 177, "all_frames", seed=5, device=dev)
    idx = store.shard(0, 1)
    batches = [idx[:B], idx[B:2 * B]]
    m = Groundlink(23, 12, 10, "all_frames").to(dev).train()
    tr = Trainer(m, opt_type="rmsprop", lr=1e-4)
    it = [0]

    def native():
        it[0] += 1
        tr.train_step(store, batches[it[0] & 1])

    out["native_trainer"] = timed(native)
    del tr, m

    m = Groundlink(23, 12, 10, "all_frames").to(dev).train()
    opt = torch.optim.RMSprop(m.parameters(), lr=1e-4)
    ev = RegressionLossEvaluator(dataset=None, split="train", device=str(dev))
    args = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                              predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))
    widths = [23, 23, 23, 3, 3, 3, 3, 36, 30, 30]
    x = dict(zip(MODEL_INPUT_ORDER, [t.contiguous() for t in torch.split(store.pack_f32(batches[0]), widths, dim=-1)]))
    labels = dict(zip(LOSS_QUANTITIES, [t.contiguous() for t in torch.split(store.labels(batches[0]), [6, 6, 6, 12], dim=-1)]))

    def module_loop():
        opt.zero_grad()
        loss = ev(x, m(x), labels, [], [], args, compute_report=False)
        loss.backward()
        opt.step()
        ev._reset_lists()

    out["module_loop"] = timed(module_loop)
    return out


def transformer_analyze_leg(dev, world: int, pk: dict, B: int = 2048, T: int = 200, n_batches: int = 6) -> dict:
    """BASELINE configs[4]: TransformerBaseline (TransformerBaseline.py:73-148; d=108, 3 heads, 3 layers, FFN 60) analysis
    pass over a stream of long windows (T=200), contiguous window shards per GPU, no collective.  142 065 600
    FLOP/window forward.  `value`: inputs resident in HBM; `e2e`: inputs in pinned host memory, H2D of
    every batch and D2H of its three outputs inside the timed region.  fp32 parameters (ctor dtype argument)."""
    from .keys import InputDataKeys as K, OutputDataKeys as O
    from .models.TransformerBaseline import TransformerBaseline
    torch.manual_seed(0)
    D = 23
    m = TransformerBaseline(D, T, dtype=torch.float32).to(dev)
    g = torch.Generator(device=dev).manual_seed(11)
    chans = [(K.POS, D), (K.VEL, D), (K.ACC, D), (K.COM_POS, 3), (K.COM_VEL, 3), (K.COM_ACC, 3)]
    x = {k: torch.randn(B, c, T, device=dev, generator=g) for k, c in chans}
    torch.set_grad_enabled(False)                     # the analysis pass (analyze.py:112-156 runs under torch.no_grad())
    for _ in range(20):                               # ~50 ms of work: a GPU that was idle needs it to reach its clocks
        m(x)
    e0, e1 = _events()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n_batches):
        out = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_batches
    kern = _per_kernel(lambda: m(x))
    torch.set_grad_enabled(True)
    kern_ms = sum(v["ms"] for v in kern.values())
    # end to end: pinned host inputs, outputs read back
    xh = {k: v.cpu().pin_memory() for k, v in x.items()}

    def e2e_pass(n):
        got = 0
        for o in m.forward_stream([xh] * n):         # pinned host inputs in, pinned host outputs out, copies overlapped
            got += 1
        assert got == n

    e2e_pass(3)
    torch.cuda.synchronize()
    e0.record()
    e2e_pass(2 * n_batches)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / (2 * n_batches)
    oh = o_last = next(iter(m.forward_stream([xh])))
    # the host link on its own: the same pinned tensors copied H2D with nothing else running (explains the e2e figure, which is
    # bound by max(compute, H2D) per batch)
    xd = {k: torch.empty_like(v, device=dev) for k, v in xh.items()}
    for k, v in xh.items():
        xd[k].copy_(v, non_blocking=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(4):
        for k, v in xh.items():
            xd[k].copy_(v, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    h2d_only_ms = e0.elapsed_time(e1) / 4
    del xd
    h2d = sum(v.numel() * v.element_size() for v in xh.values())
    d2h = sum(v.numel() * v.element_size() for v in oh.values())
    fl = 142065600.0 * B
    return {"metric": "transformer_analyze_windows_per_sec", "value": world * B / (ms * 1e-3), "unit": "windows/s",
            "ms_per_batch": ms, "batch_windows": B, "window_frames": T, "tflops": fl / (ms * 1e-3) / 1e12,
            "kernel_ms_per_batch": kern_ms, "kernels": kern,
            "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "windows/s", "ms_per_batch": ms_e2e, "h2d_bytes_per_batch": h2d,
                    "d2h_bytes_per_batch": d2h, "h2d_alone_ms_per_batch": h2d_only_ms,
                    "h2d_alone_gbs": sum(v.numel() * v.element_size() for v in xh.values()) / h2d_only_ms / 1e6,
                    "api": "TransformerBaseline.forward_stream(iterable of pinned host input dicts)"},
            "full_stream": f"2^20 windows = {(1 << 20) / (world * B / (ms * 1e-3)):.2f} s at this rate on {world} GPU(s)",
            "note": "weak scaling (contiguous window shards, no collective); d=108 rows are padded to 112 bf16 columns, heads 36->48"}
