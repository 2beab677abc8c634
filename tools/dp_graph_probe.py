"""Data-parallel small-batch steps with and without CUDA-graph replay (two graphs cut at the allreduce), N ranks on one box.
Each variant runs under its own timeout (a hung capture must not take the box down); prints, per variant, whether the
ranks agree bit for bit, the max difference to the eager run's parameters and the steady-state ms/step.
usage: python tools/dp_graph_probe.py [N=2] [kind=feedforward] [B=32]"""
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "workers", "dp_equivalence_worker.py")


def run(n, kind, B, env, port, steps=8, timed=300):
    d = tempfile.mkdtemp()
    cmd = ["timeout", "-k", "5", "150", sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, d, kind, "rmsprop", str(steps), str(B), str(timed)]
    r = subprocess.run(cmd, env=dict(os.environ, **env), capture_output=True, text=True)
    if r.returncode != 0:
        return None, f"rc {r.returncode}: {r.stderr[-1500:]}"
    return [torch.load(os.path.join(d, f"rank{i}.pt")) for i in range(n)], ""


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    kind = sys.argv[2] if len(sys.argv) > 2 else "feedforward"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    base = None
    for i, (tag, env) in enumerate((("eager, bucket/group", {"IBM_TRAIN_GRAPHS": "0", "IBM_ALLREDUCE": "overlap"}),
                                    ("eager, one allreduce", {"IBM_TRAIN_GRAPHS": "0"}),
                                    ("two graphs + allreduce", {}))):
        ranks, err = run(n, kind, B, env, 29800 + i)
        if ranks is None:
            print(f"{tag:22s} FAILED {err}")
            continue
        same = all(torch.equal(p, ranks[r]["params"][k]) for r in range(1, n) for k, p in ranks[0]["params"].items())
        if base is None:
            base = ranks[0]["params"]
        diff = max((p - base[k]).abs().max().item() for k, p in ranks[0]["params"].items())
        ms = max(r["ms_per_step"] for r in ranks)
        print(f"{tag:22s} n={n} {kind} B={B}: ranks identical {same}, max |p - eager| {diff:.3e}, {ms:.4f} ms/step "
              f"(graphs {ranks[0]['graphs']}, mode {ranks[0]['mode']}, losses {[round(x, 5) for x in ranks[0]['losses'][-2:]]})")


if __name__ == "__main__":
    main()
