"""Times the TransformerBaseline analysis pass (BASELINE configs[4]) with a per-kernel breakdown; prints JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from inferbiomechanics_b200 import bench_legs

pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
dev = torch.device("cuda", 0)
print(json.dumps(bench_legs.transformer_analyze_leg(dev, 1, pk), indent=1))

# per-shape GEMM table of one batch (M x N x K, ms, algorithmic HBM bytes / time)
from inferbiomechanics_b200 import ops  # noqa: E402
from inferbiomechanics_b200.keys import InputDataKeys as K  # noqa: E402
from inferbiomechanics_b200.models.TransformerBaseline import TransformerBaseline  # noqa: E402

m = TransformerBaseline(23, 200, dtype=torch.float32).to(dev)
x = {k: torch.randn(2048, c, 200, device=dev) for k, c in [(K.POS, 23), (K.VEL, 23), (K.ACC, 23), (K.COM_POS, 3), (K.COM_VEL, 3), (K.COM_ACC, 3)]}
for _ in range(2):
    m(x)
real = ops.gemm
recs = []


def timed(A, Bm, out, M, N, K_, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = real(A, Bm, out, M, N, K_, **kw)
    b.record()
    byts = M * A.stride(0) * 2 + M * out.stride(0) * out.element_size() + (M * kw["aux"].stride(0) * 2 if kw.get("aux") is not None else 0)
    recs.append((a, b, (M, N, K_, kw.get("act", "none"), kw.get("aux") is not None), byts))
    return r


ops.gemm = timed
m(x)
ops.gemm = real
torch.cuda.synchronize()
for a, b, shp, byts in recs:
    ms = a.elapsed_time(b)
    print(shp, round(ms * 1e3, 1), "us", round(byts / ms / 1e6), "GB/s of", round(byts / 1e6), "MB")
