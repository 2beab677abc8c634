"""Diagnostic (not a pytest): per-parameter gradient error of one denoiser training step against three CPU oracles —
fp32, bf16-mirrored forward (same ReLU gates), bf16-mirrored forward AND backward storage points.
    python tools/diag_denoiser_grads.py [small|bench]"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from oracle import loss as ol  # noqa: E402
from oracle import models as om  # noqa: E402
from oracle.gen_golden import SELECTIONS, seeded_inputs, seeded_out_labels  # noqa: E402
from oracle.seeded import seeded_state_dict  # noqa: E402

CFG = {"small": dict(B=6, F=10, d=128, heads=2, ff=256, L=2, seed=5), "bench": dict(B=8, F=50, d=512, heads=8, ff=2048, L=8, seed=7)}
ALL = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                         predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))


def main(name):
    from inferbiomechanics_b200.keys import InputDataKeys
    from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
    from inferbiomechanics_b200.models.DiffusionDenoiser import DiffusionDenoiser
    c = CFG[name]
    B, F, L, heads, seed = c["B"], c["F"], c["L"], c["heads"], c["seed"]
    m = DiffusionDenoiser(frames=F, d_model=c["d"], num_heads=heads, dim_feedforward=c["ff"], num_layers=L)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd)
    m = m.cuda()
    inputs = seeded_inputs(B, F, 23, 30, 900 + seed)
    g = torch.Generator().manual_seed(seed)
    x_t = torch.randn(B, F, 30, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    _, labels = seeded_out_labels(B, F, 901 + seed)
    sel = [list(x) for x in SELECTIONS["all"]]
    cond = om.concat_inputs(inputs)

    def run(mode, gates=None):
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        if mode == "fp32":
            x0 = om.denoiser_forward(params, cond, x_t, t, L, heads)
        else:
            x0 = om.denoiser_forward(om.bf16_weights(params), cond, x_t, t, L, heads, rnd=om.bf16_ste if mode == "fwd" else om.bf16_both,
                                     gates=gates)
        loss = ol.regression_loss(om.split30(x0), labels, *sel)["loss"]
        loss.backward()
        return {k: v.grad for k, v in params.items()}

    refs = {k: run(k) for k in ("fp32", "fwd", "both")}
    out = m({**inputs, InputDataKeys.X_T: x_t, InputDataKeys.TIMESTEP: t})
    st = m.engine().state(B, True)
    gates = [(st[f"L{l}.h"].float() > 0).float().cpu().view(B, F, -1) for l in range(L)]
    refs["gates"] = run("both", gates)
    ev = RegressionLossEvaluator(None, "train", device="cuda")
    loss = ev(inputs, out, {k: v.clone() for k, v in labels.items()}, [], [], ALL)
    loss.backward()
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
    print(f"{name}: rel-L2 of the CUDA gradient vs [fp32 | fwd-mirrored | fwd+bwd-mirrored | fwd+bwd-mirrored with the CUDA path's ReLU gates] oracle")
    for n, p in m.named_parameters():
        gg = p.grad.double().cpu()
        print(f"  {n:55s} {rel(gg, refs['fp32'][n].double()):.4f}  {rel(gg, refs['fwd'][n].double()):.4f}  "
              f"{rel(gg, refs['both'][n].double()):.4f}  {rel(gg, refs['gates'][n].double()):.4f}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small")
