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
