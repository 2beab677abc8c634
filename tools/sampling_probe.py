"""Per-kernel breakdown of one reverse-sampling (denoise) step at the BASELINE configs[3] shape (512 windows/GPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from inferbiomechanics_b200 import bench_legs, ops
from inferbiomechanics_b200.diffusion import GaussianDiffusion
from inferbiomechanics_b200.models.DiffusionDenoiser import DiffusionDenoiser

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = DiffusionDenoiser(frames=50, d_model=512, num_heads=8, dim_feedforward=2048, num_layers=8).to(dev)
eng = model.engine()
gd = GaussianDiffusion(num_timesteps=200, device=dev)
eng.xc(B, False)[:, 30:30 + eng.c_in] = torch.randn(B * 50, eng.c_in, device=dev).to(torch.bfloat16)
gd.sample(model, B, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
gd.sample(model, B, seed=1)
e1.record()
torch.cuda.synchronize()
print("graph replay ms/step", e0.elapsed_time(e1) / 200)
# per-kernel (eager, 10 steps): GEMMs by shape
real = ops.gemm
recs = []


def timed(A, Bm, out, M, N, K, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = real(A, Bm, out, M, N, K, **kw)
    b.record()
    recs.append((a, b, (M, N, K)))
    return r


ops.gemm = timed
kern = bench_legs._per_kernel(lambda: gd.sample(model, B, seed=1, steps=10, use_graph=False))
ops.gemm = real
torch.cuda.synchronize()
shapes = {}
for a, b, shp in recs:
    e = shapes.setdefault(shp, [0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b)
for shp, (n, ms) in shapes.items():
    fl = 2.0 * shp[0] * shp[1] * shp[2] * n
    print(shp, n // 10, "per step", round(ms / 10, 4), "ms/step", round(fl / ms / 1e9, 1), "TFLOP/s")
print({k: {"launches": v["launches"] // 10, "ms_per_step": round(v["ms"] / 10, 4)} for k, v in kern.items()})
