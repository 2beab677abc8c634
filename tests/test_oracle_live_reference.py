"""oracle/loss.py and oracle/models.py against the IMPORTED reference, live, on inputs the frozen fixtures never saw
(build container only: /root/reference or the oracle/_ref snapshot must be importable; skipped elsewhere).  The golden
files pin a handful of seeded cases; these draw fresh shapes, component selections and constructor arguments every case:

* RegressionLossEvaluator.__call__ (RegressionLossEvaluator.py:160-263): loss, the four component vectors, the six report
  metrics and autograd gradients — random (B, F), selections with repeats / empty lists, forces placed exactly ON the 10.0
  CoP threshold (strict >);
* the static helpers with their general contract (…:73-158): any C, C % 3, C % vec_size;
* FeedForwardBaseline (FeedForwardRegressionBaseline.py:20-121) with random hidden widths, activation, stride, output
  format and BatchNorm in eval and training mode; Groundlink (Groundlink.py:20-156) in both formats; TransformerLayer
  (TransformerBaseline.py:8-38) at random widths / heads, fp64."""
import contextlib
import io

import numpy as np
import pytest
import torch

from oracle import loss as ol
from oracle import models as om
from oracle.gen_golden import run_ref_loss, seeded_inputs
from oracle.refimport import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="needs the importable reference (build container / oracle/_ref)")


def _rand_out_labels(B, F, g, on_threshold=False):
    mk = lambda c, s=1.0: torch.randn(B, F, c, generator=g) * s
    o = {ol.COP: mk(6), ol.FORCE: mk(6, 10.0), ol.TORQUE: mk(6), ol.WRENCH: mk(12)}
    l = {ol.COP: mk(6), ol.FORCE: mk(6, 10.0), ol.TORQUE: mk(6), ol.WRENCH: mk(12)}
    if on_threshold:                                   # |force triple| == 10.0 exactly: mask must be 0 (strict >)
        l[ol.FORCE][0, 0, 0:3] = torch.tensor([6.0, 8.0, 0.0])
        l[ol.FORCE][0, 0, 3:6] = torch.tensor([0.0, 0.0, 10.0])
        if B > 1:
            l[ol.FORCE][1, -1, 0:3] = torch.tensor([6.0, 8.0, 1e-3])      # just above
    return o, l


@pytest.mark.parametrize("case", range(10))
def test_loss_call_live(case):
    ref = load_reference()
    g = torch.Generator().manual_seed(9000 + case)
    B, F = int(torch.randint(1, 9, (1,), generator=g)), int(torch.randint(1, 13, (1,), generator=g))
    pick = lambda n: [int(v) for v in torch.randint(0, n, (int(torch.randint(0, 5, (1,), generator=g)),), generator=g)]
    sel = (pick(6), pick(6), pick(6), pick(12))        # repeats and empty lists both occur
    o, l = _rand_out_labels(B, F, g, on_threshold=case % 2 == 0)
    res_ref, grads_ref = run_ref_loss(ref, o, l, sel)
    res = ol.regression_loss(o, l, *sel)
    for k in ("loss", "force", "cop", "moment", "wrench"):
        torch.testing.assert_close(res[k], torch.as_tensor(res_ref[k]), rtol=2e-6, atol=1e-7, msg=k)
    for k in ("force_report", "moment_report", "cop_report", "wrench_report", "wrench_moment_report", "com_acc_report"):
        assert abs(float(res[k]) - float(res_ref[k])) <= 2e-6 * abs(float(res_ref[k])) + 1e-7, k
    grads = ol.regression_loss_grad(o, l, *sel)
    for k in grads:
        torch.testing.assert_close(grads[k], grads_ref[k], rtol=2e-6, atol=1e-9)
    assert torch.equal(ol.mask_by_threes(l[ol.FORCE], 10.0), ref.RegressionLossEvaluator.get_mask_by_threes(l[ol.FORCE], threshold=10.0))


@pytest.mark.parametrize("C,vec", [(3, 3), (6, 3), (6, 6), (9, 3), (12, 6), (12, 4), (5, 5), (1, 1), (24, 3)])
def test_static_helpers_live(C, vec):
    R = load_reference().RegressionLossEvaluator
    g = torch.Generator().manual_seed(C * 31 + vec)
    o, l = torch.randn(4, 7, C, generator=g) * 3, torch.randn(4, 7, C, generator=g) * 3
    torch.testing.assert_close(ol.squared_diff_mean_vector(o, l), R.get_squared_diff_mean_vector(o, l), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ol.mean_norm_error(o, l, vec), R.get_mean_norm_error(o, l, vec_size=vec), rtol=1e-6, atol=1e-7)
    if C % 3 == 0:
        for thr in (0.0, 1.5, 4.0):
            assert torch.equal(ol.mask_by_threes(o, thr), R.get_mask_by_threes(o, threshold=thr))
    else:
        with pytest.raises(ValueError):
            R.get_mask_by_threes(o)
        with pytest.raises(ValueError):
            ol.mask_by_threes(o)
    if C == 6:
        torch.testing.assert_close(ol.com_acc_error(o, l), R.get_com_acc_error(o, l), rtol=1e-6, atol=1e-7)
    else:
        with pytest.raises(ValueError):
            R.get_com_acc_error(o, l)
        with pytest.raises(ValueError):
            ol.com_acc_error(o, l)
    if C % 2 == 1 and C > 1:                           # vec_size that does not divide C
        with pytest.raises(ValueError):
            R.get_mean_norm_error(o, l, vec_size=2)
        with pytest.raises(ValueError):
            ol.mean_norm_error(o, l, 2)


@pytest.mark.parametrize("case", range(8))
def test_feedforward_live(case):
    ref = load_reference()
    g = torch.Generator().manual_seed(7000 + case)
    ri = lambda a, b: int(torch.randint(a, b, (1,), generator=g))
    D, s = 23, [5, 5, 2, 10, 1, 5, 25, 5][case]
    T = s * ri(1, 6)
    act = ["sigmoid", "relu", "tanh"][case % 3]
    fmt = "all_frames" if case % 2 == 0 else "last_frame"
    bn = case in (3, 4, 5, 6)
    training = case in (5, 6)
    hidden = [ri(8, 70) for _ in range(ri(1, 4))]
    torch.manual_seed(case)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.FeedForwardBaseline(D, 2, T, fmt, act, s, 10, hidden_dims=hidden, batchnorm=bn)
    if bn:                                             # non-trivial running statistics and affine parameters
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.3)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
                mod.weight.data.copy_(1 + 0.2 * torch.randn(mod.num_features, generator=g))
                mod.bias.data.copy_(0.2 * torch.randn(mod.num_features, generator=g))
    m.train(training)
    B, F = ri(2, 7), T // s
    inputs = seeded_inputs(B, F, D, s * 3, 7100 + case)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    want = m({k: v.clone() for k, v in inputs.items()})
    new_stats = {}
    got = om.feedforward_forward(sd, inputs, act, F if fmt == "all_frames" else 1, batchnorm=bn, training=training, new_stats=new_stats)
    for k in want:
        torch.testing.assert_close(got[k], want[k].detach(), rtol=2e-5, atol=2e-6, msg=k)
    if training:                                       # the module has updated its running statistics in place
        for k, v in new_stats.items():
            torch.testing.assert_close(v, m.state_dict()[k], rtol=1e-5, atol=1e-6, msg=k)


@pytest.mark.parametrize("fmt,T,B", [("all_frames", 9, 2), ("last_frame", 23, 3), ("all_frames", 1, 1), ("last_frame", 50, 1)])
def test_groundlink_live(fmt, T, B):
    ref = load_reference()
    torch.manual_seed(T)
    m = ref.Groundlink(23, 12, 10, fmt).eval()
    inputs = seeded_inputs(B, T, 23, 30, 7300 + T)
    want = m({k: v.clone() for k, v in inputs.items()})
    got = om.groundlink_forward({k: v.clone() for k, v in m.state_dict().items()}, inputs, fmt)
    for k in want:
        torch.testing.assert_close(got[k], want[k].detach(), rtol=2e-5, atol=2e-6, msg=k)


@pytest.mark.parametrize("d,heads,ff,T,B", [(108, 3, 60, 20, 2), (64, 2, 32, 7, 3), (512, 8, 2048, 5, 1), (96, 4, 17, 33, 2), (36, 1, 8, 1, 4)])
def test_transformer_layer_live(d, heads, ff, T, B):
    ref = load_reference()
    torch.manual_seed(d + T)
    layer = ref.TransformerLayer(d, heads, ff, 0.0, dtype=torch.float64).eval() if _layer_takes_dtype(ref) else ref.TransformerLayer(d, heads, ff, 0.0).double().eval()
    x = torch.randn(B, T, d, dtype=torch.float64, requires_grad=True)
    want = layer(x)
    want.sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    got = om.transformer_layer({k: v.clone() for k, v in layer.state_dict().items()}, "", x2, heads)
    torch.testing.assert_close(got, want.detach(), rtol=1e-9, atol=1e-10)
    got.sum().backward()
    torch.testing.assert_close(x2.grad, x.grad, rtol=1e-8, atol=1e-10)


def _layer_takes_dtype(ref) -> bool:
    import inspect
    return "dtype" in inspect.signature(ref.TransformerLayer.__init__).parameters


@pytest.mark.parametrize("dofs,T,emb,layers,heads,ffw,B", [(23, 20, 30, 3, 3, 60, 2), (10, 9, 9, 1, 2, 16, 3), (23, 64, 30, 2, 6, 24, 1),
                                                          (5, 3, 6, 4, 1, 8, 2)])
def test_transformer_baseline_live(dofs, T, emb, layers, heads, ffw, B):
    """TransformerBaseline.forward (TransformerBaseline.py:104-148) composed from the reference's own sub-modules (its forward
    needs key constants that do not exist) with non-default constructor arguments; outputs and the gradient of
    every parameter (fp64)."""
    ref = load_reference()
    torch.manual_seed(dofs * 100 + T)
    m = ref.TransformerBaseline(dofs, T, temporal_embedding_dim=emb, num_layers=layers, num_heads=heads, dim_feedforward=ffw).eval()
    x = {k: torch.randn(B, c, T, dtype=torch.float64) for k, c in (("pos", dofs), ("vel", dofs), ("acc", dofs), ("comPos", 3),
                                                                     ("comVel", 3), ("comAcc", 3))}
    vecs = torch.cat([x["pos"], x["vel"], x["acc"], x["comPos"], x["comVel"], x["comAcc"]], dim=1).transpose(1, 2)
    e = m.temporal_embedding(torch.arange(vecs.size(1))).expand(B, T, m.temporal_embedding_dim)
    vecs = torch.cat([vecs, e], dim=2)
    for layer in m.transformer_layers:
        vecs = layer(vecs)
    output = m.fc(vecs)
    want = {"contact": m.contact_sigmoid(output[:, :, :2]).transpose(1, 2),
            "comAcc": m.com_attention(vecs, vecs, x["comAcc"].transpose(1, 2)).transpose(1, 2),
            "contactForces": output[:, :, 5:].transpose(1, 2)}
    w = {k: torch.randn(v.shape, dtype=torch.float64) for k, v in want.items()}
    sum((want[k] * w[k]).sum() for k in want).backward()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    got = om.transformer_forward(sd, x, layers, heads)
    for k in want:
        torch.testing.assert_close(got[k], want[k].detach(), rtol=1e-9, atol=1e-10, msg=k)
    sum((got[k] * w[k]).sum() for k in got).backward()
    for n, p in m.named_parameters():
        ref_g = p.grad if p.grad is not None else torch.zeros_like(p)
        got_g = sd[n].grad if sd[n].grad is not None else torch.zeros_like(p)
        torch.testing.assert_close(got_g, ref_g, rtol=1e-7, atol=1e-10, msg=n)
