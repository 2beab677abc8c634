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


def test_groundlink_cnn_dropout_matches_masked_emulation():
    """cnn_dropout > 0 in training mode (nn.Dropout in front of every Conv1d, Groundlink.py:41): the four Philox masks are
    regenerated through the same C-ABI call over identically shaped buffers and the drop-in's forward / backward is
    compared with a plain fp32 torch emulation using them (dropout -> replicate-pad conv -> ELU, x4; then the MLP with
    fc_dropout = 0).  Tolerances as for the golden Groundlink test (bf16 operands through 7 layers)."""
    import torch.nn.functional as Fn
    from inferbiomechanics_b200 import ops
    from inferbiomechanics_b200.models.Groundlink import Groundlink
    B, T, p = 3, 20, 0.25
    torch.manual_seed(11)
    m = Groundlink(23, 12, 10, "all_frames", cnn_dropout=p, fc_dropout=0.0).cuda().train()
    inputs = seeded_inputs(B, T, 23, 30, 21)
    out = m(inputs)
    y = torch.cat([out[k] for k in (ol.COP, ol.FORCE, ol.TORQUE, ol.WRENCH)], dim=-1)          # (B, T, 30)
    gy = seeded_out_labels(B, T, 31)[0]
    gy = torch.cat([gy[k] for k in (ol.COP, ol.FORCE, ol.TORQUE, ol.WRENCH)], dim=-1).cuda()
    for q in m.parameters():
        q.grad = None
    (y * gy).sum().backward()
    eng = m.engine()
    st, Tp, Mp = eng._state(B, T)
    slack = st["slack"]
    masks = []
    for i in range(4):
        ones = torch.ones_like(st[f"x{i}_full"])
        mk = torch.empty_like(ones)
        ops.dropout(ones, mk, p, eng.CNN_DROPOUT_SEED, 4 * eng.step + i)
        mk = mk[slack:slack + Mp].view(B, Tp, -1)[:, 3:3 + T, :eng.ch[i]].float()
        assert abs((mk == 0).float().mean().item() - p) < 0.05
        masks.append(mk)
    x = st["x0"].view(B, Tp, -1)[:, 3:3 + T, :eng.ch[0]].float()                               # the packed (bf16-rounded) input frames
    params = {n: q.detach().float().clone().requires_grad_(True) for n, q in m.named_parameters()}
    h = x
    for i, pos in enumerate((1, 4, 7, 10)):
        h = (h * masks[i]).transpose(1, 2)                                                     # (B, C, T)
        h = Fn.conv1d(Fn.pad(h, (3, 3), mode="replicate"), params[f"cnn.{pos}.weight"], params[f"cnn.{pos}.bias"])
        h = Fn.elu(h).transpose(1, 2)
    h = Fn.elu(h @ params["fc.2.weight"].t() + params["fc.2.bias"])
    h = Fn.elu(h @ params["fc.5.weight"].t() + params["fc.5.bias"])
    ref = h @ params["fc.8.weight"].t()
    err = (y.detach() - ref.detach()).abs().max().item()
    assert err <= 3e-2 * ref.abs().max().item(), f"forward: {err} vs {ref.abs().max().item()}"
    (ref * gy).sum().backward()
    for n, q in m.named_parameters():
        got, want = q.grad.double().reshape(-1).cpu(), params[n].grad.double().reshape(-1).cpu()
        rel = (got - want).norm().item() / (want.norm().item() + 1e-12)
        cos = torch.dot(got, want).item() / (got.norm().item() * want.norm().item() + 1e-30)
        assert rel <= 0.12 and cos >= 0.985, f"{n}: rel L2 {rel:.4f}, cosine {cos:.4f}"
    # eval mode ignores the dropout; a second training forward draws new masks
    m.eval()
    e1 = m(inputs)[ol.FORCE].clone()
    m.train()
    t2 = m(inputs)[ol.FORCE]
    assert not torch.equal(t2, out[ol.FORCE]) and not torch.equal(e1, t2)


@pytest.mark.parametrize("name,fmt", [("gl_k5_d2", "all_frames"), ("gl_k3_d4", "last_frame"), ("gl_k9_d1", "all_frames")])
def test_groundlink_constructor_variants_match_reference_golden(golden, name, fmt):
    """Groundlink(cnn_kernel, fc_depth) other than the defaults (Groundlink.py:20,41,51-62), golden from the imported reference
    (tests/golden/ctor_variants.npz): outputs, loss and every parameter gradient at the bars of the default model."""
    from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
    from inferbiomechanics_b200.models.Groundlink import Groundlink
    g = golden("ctor_variants.npz")
    D, J, H, B, T, seed, iseed, lseed, k, depth = (int(v) for v in g[f"{name}/meta"])
    m = Groundlink(D, J, H, fmt, cnn_kernel=k, fc_depth=depth)
    m.load_state_dict(seeded_state_dict({kk: tuple(v.shape) for kk, v in m.state_dict().items()}, seed))
    m = m.cuda().eval()
    inputs = seeded_inputs(B, T, D, H * 3, iseed)
    _, labels = seeded_out_labels(B, T if fmt == "all_frames" else 1, lseed)
    out = m(inputs)
    for q in Q:
        ref = g[f"{name}/out/{q}"]
        assert tuple(out[q].shape) == ref.shape
        assert np.abs(out[q].detach().cpu().numpy() - ref).max() <= 3e-2 * np.abs(ref).max(), q
    ev = RegressionLossEvaluator(None, "train", device="cuda")
    loss = ev(inputs, out, {kk: v.clone() for kk, v in labels.items()}, [], [], ALL)
    np.testing.assert_allclose(loss.item(), float(g[f"{name}/loss"]), rtol=2e-2)
    loss.backward()
    for n, p in m.named_parameters():
        got = strided_sample(p.grad).double().cpu()
        ref = torch.from_numpy(g[f"{name}/grad_sample/{n}"]).double()
        rel = (got - ref).norm().item() / (ref.norm().item() + 1e-12)
        cos = torch.dot(got, ref).item() / (got.norm().item() * ref.norm().item() + 1e-30)
        assert rel <= 0.12 and cos >= 0.985, f"{n}: rel L2 {rel:.4f}, cosine {cos:.4f}"
    with pytest.raises(NotImplementedError):
        Groundlink(D, J, H, fmt, cnn_kernel=6)


def test_batch_of_one_latency_path_equals_eager(monkeypatch):
    """the viewers' ``model(inputs)`` on ONE window under no_grad replays a captured CUDA graph (pack + GEMMs);
    same bits as the eager launches, picks up weight changes (the graph reads the arena through stable pointers; re-laid-out
    conv weights are refreshed outside it), leaves the training path alone."""
    from inferbiomechanics_b200.models.FeedForwardRegressionBaseline import FeedForwardBaseline
    from inferbiomechanics_b200.models.Groundlink import Groundlink
    torch.manual_seed(0)
    cases = [(FeedForwardBaseline(23, 2, 50, "all_frames", "sigmoid", 5, 10).cuda().eval(), 10, 15),
             (Groundlink(23, 12, 10, "last_frame").cuda().eval(), 50, 30)]
    for m, F, hist in cases:
        xs = [seeded_inputs(1, F, 23, hist, 40 + i) for i in range(5)]
        with torch.no_grad():
            monkeypatch.setenv("IBM_INFER_GRAPHS", "0")
            eager = [{k: v.clone() for k, v in m(x).items()} for x in xs]
            monkeypatch.setenv("IBM_INFER_GRAPHS", "1")
            graphed = [{k: v.clone() for k, v in m(x).items()} for x in xs]      # calls 1-2 eager, 3 captures, 4-5 replay
            assert any(st["graph"] is not None for st in m._lat.values())
            for a, b in zip(eager, graphed):
                for k in Q:
                    assert torch.equal(a[k], b[k]), k
            # weights change under the captured graph: the replay must see them
            with torch.no_grad():
                for p in m.parameters():
                    p.mul_(0.5)
            monkeypatch.setenv("IBM_INFER_GRAPHS", "0")
            want = {k: v.clone() for k, v in m(xs[0]).items()}
            monkeypatch.setenv("IBM_INFER_GRAPHS", "1")
            got = m(xs[0])
            for k in Q:
                assert torch.equal(want[k], got[k]) and not torch.equal(want[k], eager[0][k])
        out = m(xs[0])                                       # grad enabled: the autograd path, not the graph
        assert out[Q[0]].requires_grad
