"""GPU parity of the drop-in Groundlink (implicit-GEMM temporal CNN + per-frame MLP) against golden vectors
from the imported reference (/root/reference/src/models/Groundlink.py), eval mode (dropout off) like the fixture.

Tolerances: outputs |err| <= 3e-2 * max|ref| (7 bf16 GEMM layers deep); loss rtol 2e-2; parameter gradients
relative L2 <= 0.12 per tensor with cosine >= 0.985 (bf16 activations and bf16 upstream gradients)."""
import argparse

import numpy as np
import pytest
import torch

from oracle import loss as ol
from oracle.gen_golden import seeded_inputs, seeded_out_labels
from oracle.seeded import seeded_state_dict, strided_sample

pytestmark = pytest.mark.gpu
Q = (ol.COP, ol.FORCE, ol.TORQUE, ol.WRENCH)
ALL = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                         predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))


@pytest.mark.parametrize("name,fmt", [("all_t50", "all_frames"), ("last_t20", "last_frame")])
def test_groundlink_matches_reference_golden(golden, name, fmt):
    from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
    from inferbiomechanics_b200.models.Groundlink import Groundlink
    g = golden("groundlink.npz")
    D, J, H, B, T, seed, iseed, lseed = (int(v) for v in g[f"{name}/meta"])
    m = Groundlink(D, J, H, fmt)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd)                                   # reference state_dict keys
    m = m.cuda().eval()
    inputs = seeded_inputs(B, T, D, H * 3, iseed)
    _, labels = seeded_out_labels(B, T if fmt == "all_frames" else 1, lseed)
    out = m(inputs)
    for k in Q:
        ref = g[f"{name}/out/{k}"]
        assert tuple(out[k].shape) == ref.shape
        err = np.abs(out[k].detach().cpu().numpy() - ref).max()
        assert err <= 3e-2 * np.abs(ref).max(), f"{k}: {err} vs {np.abs(ref).max()}"
    ev = RegressionLossEvaluator(None, "train", device="cuda")
    loss = ev(inputs, out, {k: v.clone() for k, v in labels.items()}, [], [], ALL)
    np.testing.assert_allclose(loss.item(), float(g[f"{name}/loss"]), rtol=2e-2)
    for p in m.parameters():
        p.grad = None
    loss.backward()
    for n, p in m.named_parameters():
        got = strided_sample(p.grad).double().cpu()
        ref = torch.from_numpy(g[f"{name}/grad_sample/{n}"]).double()
        rel = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
        cos = torch.dot(got, ref).item() / (got.norm().item() * ref.norm().item() + 1e-30)
        assert rel <= 0.12 and cos >= 0.985, f"{n}: rel L2 {rel:.4f}, cosine {cos:.4f}"


def test_groundlink_init_and_dropout_training_step():
    """Same init rule as the reference (Xavier-normal/ReLU gain before ELUs, zero bias), and a training-mode
    forward/backward with Dropout(0.2) active runs and is reproducible for a fixed step counter."""
    from inferbiomechanics_b200.models.Groundlink import Groundlink
    torch.manual_seed(0)
    m = Groundlink(23, 12, 10, "all_frames")
    assert float(m.cnn[1].bias.abs().sum()) == 0.0 and float(m.fc[2].bias.abs().sum()) == 0.0
    std = m.cnn[4].weight.std().item()
    want = (2.0 ** 0.5) * (2.0 / ((128 + 128) * 7)) ** 0.5
    assert abs(std - want) / want < 0.05
    m = m.cuda().train()
    inputs = seeded_inputs(4, 50, 23, 30, 7)
    out = m(inputs)
    tot = sum(v.float().sum() for v in out.values())
    for p in m.parameters():
        p.grad = None
    tot.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    m.eval()
    a = m(inputs)[ol.FORCE].clone()
    b = m(inputs)[ol.FORCE].clone()
    assert torch.equal(a, b)
