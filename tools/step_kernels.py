#!/usr/bin/env python
"""Per-kernel GPU time inside the denoiser training step (BASELINE configs[1]: 4096 windows x 50 frames, d = 512, 8 layers),
CUDA events around every C-ABI launch of instrumented steps, median of `reps` steps after `warm` plain steps.  A/B aid: run it
under different environment switches in ONE gpurun call (boxes differ by +-4 %).

    python tools/step_kernels.py [reps=5] [warm=8]
"""
import statistics
import sys

import torch

sys.path.insert(0, ".")
from inferbiomechanics_b200.data.window_store import WindowStore  # noqa: E402
from inferbiomechanics_b200.diffusion import GaussianDiffusion  # noqa: E402
from inferbiomechanics_b200.models.DiffusionDenoiser import DiffusionDenoiser  # noqa: E402
from inferbiomechanics_b200.trainer import Trainer  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    B, F = 4096, 50
    model = DiffusionDenoiser(frames=F, d_model=512, num_heads=8, dim_feedforward=2048, num_layers=8).to(dev)
    tr = Trainer(model, opt_type="rmsprop", lr=1e-4, diffusion=GaussianDiffusion(device=dev), seed=1234)
    store = WindowStore.synthetic(B * 2, F, 1, 177, "all_frames", seed=1234, device=dev)
    idx = store.shard(0, 1)
    batches = [idx[:B], idx[B:2 * B]]
    for i in range(warm):
        tr.train_step(store, batches[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        tr.train_step(store, batches[i % 2])
    e1.record()
    torch.cuda.synchronize()
    print(f"step {e0.elapsed_time(e1) / 10:.3f} ms ({B / (e0.elapsed_time(e1) / 10) * 1e3:.0f} windows/s)")
    per, gemm = {}, {}
    for r in range(reps):
        tr.profile_gemms(store, batches[r % 2])
        for k, v in tr.last_kernel_ms.items():
            per.setdefault(k, []).append((v["ms"], v["launches"]))
        shapes = {}
        for ms, fl, shp in tr.last_gemm_records:
            s = shapes.setdefault(shp, [0.0, 0])
            s[0] += ms
            s[1] += 1
        for shp, (ms, n) in shapes.items():
            gemm.setdefault(shp, []).append((ms, n))
    print("kernel                        launches   ms/step   us/launch")
    for k, v in sorted(per.items(), key=lambda kv: -statistics.median(x[0] for x in kv[1])):
        ms = statistics.median(x[0] for x in v)
        if ms < 0.05:
            continue
        print(f"{k:30s} {v[0][1]:6d} {ms:10.3f} {ms / v[0][1] * 1e3:10.1f}")
    tot = 0.0
    for shp, v in sorted(gemm.items(), key=lambda kv: -statistics.median(x[0] for x in kv[1])):
        ms = statistics.median(x[0] for x in v)
        tot += ms
        if ms < 0.2:
            continue
        M, N, K = shp
        print(f"gemm {M}x{N}x{K:<22d} {v[0][1]:4d} {ms:10.3f} {ms / v[0][1] * 1e3:10.1f}   {2.0 * M * N * K * v[0][1] / ms / 1e9:7.1f} TFLOP/s")
    print(f"gemm total {tot:.3f} ms")


if __name__ == "__main__":
    main()
