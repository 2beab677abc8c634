"""Per-kernel breakdown of one native Groundlink training step (B windows x T=50 frames)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from inferbiomechanics_b200 import bench_legs, ops
from inferbiomechanics_b200.data.window_store import WindowStore
from inferbiomechanics_b200.models.Groundlink import Groundlink
from inferbiomechanics_b200.trainer import Trainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
store = WindowStore.synthetic(B, 50, 1, 177, "all_frames", seed=5, device=dev)
idx = store.shard(0, 1)[:B]
m = Groundlink(23, 12, 10, "all_frames").to(dev).train()
tr = Trainer(m, opt_type="rmsprop", lr=1e-4)
for _ in range(3):
    tr.train_step(store, idx)
real = ops.gemm
recs = []


def timed(A, Bm, out, M, N, K, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = real(A, Bm, out, M, N, K, **kw)
    b.record()
    recs.append((a, b, (M, N, K, kw.get("taps", 1), bool(kw.get("accumulate", False)))))
    return r


ops.gemm = timed
kern = bench_legs._per_kernel(lambda: tr.train_step(store, idx))
ops.gemm = real
torch.cuda.synchronize()
shapes = {}
for a, b, shp in recs:
    e = shapes.setdefault(shp, [0, 0.0])
    e[0] += 1
    e[1] += a.elapsed_time(b)
tot = 0
for shp, (n, ms) in sorted(shapes.items(), key=lambda kv: -kv[1][1]):
    fl = 2.0 * shp[0] * shp[1] * shp[2] * n
    tot += ms
    print(shp, n, "launches", round(ms, 4), "ms", round(fl / ms / 1e9, 1), "TFLOP/s")
print("gemm total ms", round(tot, 3))
print(json.dumps(kern, indent=0))
