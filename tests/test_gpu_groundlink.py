"""GPU parity of the drop-in Groundlink (implicit-GEMM temporal CNN + per-frame MLP) against golden vectors
from the imported reference (/root/reference/src/models/Groundlink.py), eval mode (dropout off) like the fixture.

Tolerances: outputs |err| <= 3e-2 * max|ref| (7 bf16 GEMM layers deep); loss rtol 2e-2; parameter gradients
relative L2 <= 0.12 per tensor with cosine >= 0.985 (bf16 activations and bf16 upstream gradients)."""
import argparse
import math

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


@pytest.mark.parametrize("fmt", ["all_frames", "last_frame"])
def test_groundlink_native_trainer_matches_module_loop(fmt):
    """Trainer on a Groundlink (window store -> padded-row packer -> engine forward -> fused loss on the strided output
    view -> engine backward into the flat arena -> fused SGD) against the reference's loop shape on the SAME weights and
    windows (module forward -> RegressionLossEvaluator -> loss.backward() -> torch.optim.SGD.step()).  Both run the same
    kernels, so losses agree to fp32 rounding and the parameters after 3 steps to bf16-shadow rounding.  Dropout off
    (fc_dropout=0) so the two runs see the same network; host-fed step included."""
    import argparse
    from inferbiomechanics_b200.data.window_store import WindowStore
    from inferbiomechanics_b200.keys import LOSS_QUANTITIES, MODEL_INPUT_ORDER
    from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
    from inferbiomechanics_b200.models.Groundlink import Groundlink
    from inferbiomechanics_b200.trainer import Trainer
    T, B, lr = 20, 8, 1e-3
    store = WindowStore.synthetic(64, T, 1, 177, fmt, seed=9, trial_len=120)
    torch.manual_seed(5)
    a = Groundlink(23, 12, 10, fmt, fc_dropout=0.0).cuda().train()
    b = Groundlink(23, 12, 10, fmt, fc_dropout=0.0).cuda().train()
    b.load_state_dict(a.state_dict())
    tr = Trainer(a, opt_type="sgd", lr=lr)
    opt = torch.optim.SGD(b.parameters(), lr=lr)
    ev = RegressionLossEvaluator(dataset=None, split="train", device="cuda")
    args = argparse.Namespace(predict_grf_components=list(range(6)), predict_cop_components=list(range(6)),
                              predict_moment_components=list(range(6)), predict_wrench_components=list(range(12)))
    widths = [23, 23, 23, 3, 3, 3, 3, 36, 30, 30]
    idx_all = store.shard(0, 1)
    for step in range(3):
        idx = idx_all[step * B:(step + 1) * B]
        res = tr.train_step(store, idx)
        x = store.pack_f32(idx)                                          # (B, T, 177) fp32, model concat order
        inputs = dict(zip(MODEL_INPUT_ORDER, torch.split(x, widths, dim=-1)))
        lab = store.labels(idx)
        labels = dict(zip(LOSS_QUANTITIES, torch.split(lab, [6, 6, 6, 12], dim=-1)))
        opt.zero_grad()
        loss = ev(inputs, b(inputs), labels, [], [], args)
        loss.backward()
        opt.step()
        np.testing.assert_allclose(res[0].item(), loss.item(), rtol=2e-3)
    for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        d0 = (p - q).abs().max().item()
        assert d0 <= 2e-2 * lr * 50 + 1e-6 * q.abs().max().item(), f"{n}: parameters diverged by {d0}"
    # host-fed step (pinned CPU tensors -> one packing kernel -> the same native step)
    hb = tr.make_host_batch(B, seed=3, frames=T)
    l0 = tr.train_step_host(hb["inputs"], hb["labels"])
    assert math.isfinite(l0)
    # evaluation pass
    r = tr.eval_step(store, idx_all[:B])
    assert torch.isfinite(r[0])
