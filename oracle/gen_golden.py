"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference).
TEST INFRASTRUCTURE ONLY.  Run in the build container:  ``python -m oracle.gen_golden``

The reference cannot travel to the GPU box, so its outputs on seeded inputs are frozen here as
small fixtures; weights are rebuilt from seeds by ``oracle/seeded.py`` on both sides.

Fixtures
--------
loss_call.npz          RegressionLossEvaluator.__call__ (src/loss/RegressionLossEvaluator.py:160-322)
                       on seeded outputs/labels for several component selections, plus autograd
                       gradients of the loss w.r.t. the outputs, plus the 4 static helpers.
ff.npz                 FeedForwardBaseline (src/models/FeedForwardRegressionBaseline.py) forward,
                       loss and parameter gradients; sigmoid/relu/tanh; all_frames/last_frame;
                       batchnorm (eval).
ff_bn_train.npz        FeedForwardBaseline with batchnorm=True in training mode: outputs, loss, parameter
                       gradients and the updated running statistics after one forward.
groundlink.npz         Groundlink (src/models/Groundlink.py) forward + loss + gradients.
transformer.npz        TransformerLayer stack + heads composed exactly as
                       TransformerBaseline.forward (src/models/TransformerBaseline.py:104-148), fp64.
denoiser_layers.npz    the reference TransformerLayer at the denoiser's width (d=64/128 test sizes,
                       fp32) — pins the layer the builder-owned denoiser re-uses.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import loss as oloss
from .refimport import load_reference
from .seeded import seeded_state_dict, seeded_tensor, strided_sample
from .windows import input_widths

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def ns_args(grf, cop, moment, wrench):
    return argparse.Namespace(predict_grf_components=list(grf), predict_cop_components=list(cop),
                              predict_moment_components=list(moment), predict_wrench_components=list(wrench))


SELECTIONS = {
    "all": (range(6), range(6), range(6), range(12)),
    "analyze_default": ([1], [], [], []),
    "repeat": ([1, 1, 4], [0, 5], [2], [11, 3, 3]),
}


def seeded_out_labels(B, F, seed):
    o = {oloss.COP: seeded_tensor((B, F, 6), seed + 1), oloss.FORCE: seeded_tensor((B, F, 6), seed + 2, 10.0),
         oloss.TORQUE: seeded_tensor((B, F, 6), seed + 3), oloss.WRENCH: seeded_tensor((B, F, 12), seed + 4)}
    l = {oloss.COP: seeded_tensor((B, F, 6), seed + 5), oloss.FORCE: seeded_tensor((B, F, 6), seed + 6, 10.0),
         oloss.TORQUE: seeded_tensor((B, F, 6), seed + 7), oloss.WRENCH: seeded_tensor((B, F, 12), seed + 8)}
    return o, l


def seeded_inputs(B, F, D, hist_cols, seed):
    return {k: seeded_tensor((B, F, c), seed + 17 * i) for i, (k, c) in enumerate(input_widths(D, hist_cols).items())}


def run_ref_loss(ref, outputs, labels, sel):
    ev = ref.RegressionLossEvaluator(dataset=None, split="dev")
    outs = {k: v.clone().requires_grad_(True) for k, v in outputs.items()}
    loss = ev(None, dict(outs), {k: v.clone() for k, v in labels.items()}, [], [], ns_args(*sel))
    loss.backward()
    res = dict(loss=loss.detach(), force=ev.force_losses[-1].detach(), cop=ev.cop_losses[-1].detach(),
               moment=ev.moment_losses[-1].detach(), wrench=ev.wrench_losses[-1].detach(),
               force_report=ev.force_reported_metrics[-1], moment_report=ev.moment_reported_metrics[-1],
               cop_report=ev.cop_reported_metrics[-1], wrench_report=ev.wrench_reported_metrics[-1],
               wrench_moment_report=ev.wrench_moment_reported_metrics[-1],
               com_acc_report=ev.com_acc_reported_metrics[-1])
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in outs.items()}
    return res, grads


def gen_loss(ref):
    d = {}
    for case, (B, F, seed) in {"b4f10": (4, 10, 100), "b3f1": (3, 1, 200), "b7f50": (7, 50, 300)}.items():
        o, l = seeded_out_labels(B, F, seed)
        for sname, sel in SELECTIONS.items():
            res, grads = run_ref_loss(ref, o, l, sel)
            for k, v in res.items():
                d[f"{case}/{sname}/{k}"] = np.asarray(v, dtype=np.float64 if isinstance(v, float) else None)
            for k, v in grads.items():
                d[f"{case}/{sname}/grad/{k}"] = v.numpy()
        R = ref.RegressionLossEvaluator
        d[f"{case}/mask10"] = R.get_mask_by_threes(l[oloss.FORCE], threshold=10.0).numpy()
        d[f"{case}/sqdiff_wrench"] = R.get_squared_diff_mean_vector(o[oloss.WRENCH], l[oloss.WRENCH]).numpy()
        d[f"{case}/mne6_wrench"] = R.get_mean_norm_error(o[oloss.WRENCH], l[oloss.WRENCH], vec_size=6).numpy()
        d[f"{case}/com_acc"] = R.get_com_acc_error(o[oloss.FORCE], l[oloss.FORCE]).numpy()
        d[f"{case}/meta"] = np.array([B, F, seed])
    np.savez_compressed(os.path.join(OUT, "loss_call.npz"), **d)


def grads_summary(model, d, prefix):
    for n, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        d[f"{prefix}/grad_sample/{n}"] = strided_sample(g).numpy()
        d[f"{prefix}/grad_sum/{n}"] = np.array([g.double().sum().item(), g.double().abs().sum().item()])


def gen_ff(ref):
    d = {}
    D, T, s, B = 23, 50, 5, 6
    cases = {
        "sigmoid_all": dict(act="sigmoid", fmt="all_frames", hidden=[64, 48], bn=False),
        "relu_last": dict(act="relu", fmt="last_frame", hidden=[32], bn=False),
        "tanh_all": dict(act="tanh", fmt="all_frames", hidden=[40, 40, 24], bn=False),
        "sigmoid_bn": dict(act="sigmoid", fmt="all_frames", hidden=[64, 48], bn=True),
    }
    for ci, (name, c) in enumerate(cases.items()):
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            m = ref.FeedForwardBaseline(D, 2, T, c["fmt"], c["act"], s, 10, hidden_dims=c["hidden"], batchnorm=c["bn"])
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        seed = 1000 + ci
        m.load_state_dict(seeded_state_dict(shapes, seed))
        m.eval()
        F = T // s
        inputs = seeded_inputs(B, F, D, s * 3, 2000 + ci)
        Fo = F if c["fmt"] == "all_frames" else 1
        _, labels = seeded_out_labels(B, Fo, 3000 + ci)
        out = m({k: v.clone() for k, v in inputs.items()})
        ev = ref.RegressionLossEvaluator(dataset=None, split="train")
        loss = ev(None, dict(out), {k: v.clone() for k, v in labels.items()}, [], [], ns_args(*SELECTIONS["all"]))
        loss.backward()
        for k, v in out.items():
            d[f"{name}/out/{k}"] = v.detach().numpy()
        d[f"{name}/loss"] = loss.detach().numpy()
        grads_summary(m, d, name)
        d[f"{name}/meta"] = np.array([D, T, s, B, seed, 2000 + ci, 3000 + ci])
        d[f"{name}/hidden"] = np.array(c["hidden"])
    np.savez_compressed(os.path.join(OUT, "ff.npz"), **d)


def gen_ff_bn_train(ref):
    """FeedForwardBaseline with batchnorm=True in TRAINING mode (batch statistics, running-stat update):
    one forward + loss + backward of the reference on seeded inputs (FeedForward…py:68-77)."""
    d = {}
    D, T, s = 23, 50, 5
    for ci, (name, act, hidden, B) in enumerate([("sigmoid_b16", "sigmoid", [64, 48], 16), ("relu_b300", "relu", [96], 300)]):
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            m = ref.FeedForwardBaseline(D, 2, T, "all_frames", act, s, 10, hidden_dims=hidden, batchnorm=True)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        seed = 1500 + ci
        m.load_state_dict(seeded_state_dict(shapes, seed))
        m.train()
        F = T // s
        inputs = seeded_inputs(B, F, D, s * 3, 2500 + ci)
        _, labels = seeded_out_labels(B, F, 3500 + ci)
        out = m({k: v.clone() for k, v in inputs.items()})
        ev = ref.RegressionLossEvaluator(dataset=None, split="train")
        loss = ev(None, dict(out), {k: v.clone() for k, v in labels.items()}, [], [], ns_args(*SELECTIONS["all"]))
        loss.backward()
        for k, v in out.items():
            d[f"{name}/out/{k}"] = v.detach().numpy()
        d[f"{name}/loss"] = loss.detach().numpy()
        grads_summary(m, d, name)
        for k, v in m.state_dict().items():
            if "running_" in k:
                d[f"{name}/buffer/{k}"] = v.numpy()
            if k.endswith("num_batches_tracked"):
                d[f"{name}/buffer/{k}"] = np.array(int(v))
        d[f"{name}/meta"] = np.array([D, T, s, B, seed, 2500 + ci, 3500 + ci])
        d[f"{name}/hidden"] = np.array(hidden)
    np.savez_compressed(os.path.join(OUT, "ff_bn_train.npz"), **d)


def gen_groundlink(ref):
    d = {}
    D, J, H = 23, 12, 10
    for ci, (name, fmt, B, T) in enumerate([("all_t50", "all_frames", 3, 50), ("last_t20", "last_frame", 2, 20)]):
        m = ref.Groundlink(D, J, H, fmt)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        seed = 1100 + ci
        m.load_state_dict(seeded_state_dict(shapes, seed))
        m.eval()
        inputs = seeded_inputs(B, T, D, H * 3, 2100 + ci)
        Fo = T if fmt == "all_frames" else 1
        _, labels = seeded_out_labels(B, Fo, 3100 + ci)
        out = m({k: v.clone() for k, v in inputs.items()})
        ev = ref.RegressionLossEvaluator(dataset=None, split="train")
        loss = ev(None, dict(out), {k: v.clone() for k, v in labels.items()}, [], [], ns_args(*SELECTIONS["all"]))
        loss.backward()
        for k, v in out.items():
            d[f"{name}/out/{k}"] = v.detach().numpy()
        d[f"{name}/loss"] = loss.detach().numpy()
        grads_summary(m, d, name)
        d[f"{name}/meta"] = np.array([D, J, H, B, T, seed, 2100 + ci, 3100 + ci])
    np.savez_compressed(os.path.join(OUT, "groundlink.npz"), **d)


def gen_transformer(ref):
    """TransformerBaseline.forward composed by hand from the reference sub-modules (the class's own
    forward needs key constants that do not exist) — lines 104-148, fp64."""
    d = {}
    D = 23
    for ci, (name, B, T) in enumerate([("t20", 2, 20), ("t200", 1, 200)]):
        m = ref.TransformerBaseline(D, T)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        seed = 1200 + ci
        m.load_state_dict(seeded_state_dict(shapes, seed, dtype=torch.float64))
        m.eval()
        x = {k: seeded_tensor((B, c, T), 2200 + ci + 10 * i, dtype=torch.float64)
             for i, (k, c) in enumerate([("pos", D), ("vel", D), ("acc", D), ("comPos", 3), ("comVel", 3), ("comAcc", 3)])}
        with torch.no_grad():
            vecs = torch.cat([x["pos"], x["vel"], x["acc"], x["comPos"], x["comVel"], x["comAcc"]], dim=1).transpose(1, 2)
            emb = m.temporal_embedding(torch.arange(vecs.size(1))).expand(B, T, m.temporal_embedding_dim)
            vecs = torch.cat([vecs, emb], dim=2)
            per_layer = []
            for layer in m.transformer_layers:
                vecs = layer(vecs)
                per_layer.append(vecs.clone())
            output = m.fc(vecs)
            blend = m.com_attention(vecs, vecs, x["comAcc"].transpose(1, 2))
            d[f"{name}/contact"] = m.contact_sigmoid(output[:, :, :2]).transpose(1, 2).numpy()
            d[f"{name}/comAcc"] = blend.transpose(1, 2).numpy()
            d[f"{name}/contactForces"] = output[:, :, 5:].transpose(1, 2).numpy()
            d[f"{name}/layer0"] = per_layer[0].numpy()
            d[f"{name}/layer_last"] = per_layer[-1].numpy()
        d[f"{name}/meta"] = np.array([D, B, T, seed, 2200 + ci])
    np.savez_compressed(os.path.join(OUT, "transformer.npz"), **d)


def gen_denoiser_layers(ref):
    """Reference TransformerLayer at the denoiser's configuration family (fp32, heads×64)."""
    d = {}
    for ci, (name, dm, heads, ff, B, T) in enumerate([("d128", 128, 2, 256, 3, 50), ("d512", 512, 8, 2048, 1, 50)]):
        layer = ref.TransformerLayer(dm, heads, ff, 0.0, dtype=torch.float32)
        shapes = {k: tuple(v.shape) for k, v in layer.state_dict().items()}
        seed = 1300 + ci
        layer.load_state_dict(seeded_state_dict(shapes, seed))
        x = seeded_tensor((B, T, dm), 2300 + ci).requires_grad_(True)
        y = layer(x)
        gy = seeded_tensor((B, T, dm), 2400 + ci)
        (y * gy).sum().backward()
        d[f"{name}/y"] = y.detach().numpy()
        d[f"{name}/dx"] = x.grad.numpy()
        grads_summary(layer, d, name)
        d[f"{name}/meta"] = np.array([dm, heads, ff, B, T, seed, 2300 + ci, 2400 + ci])
    np.savez_compressed(os.path.join(OUT, "denoiser_layers.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)      # deterministic reduction order for the frozen vectors
    ref = load_reference()
    if "--only-bn-train" in __import__("sys").argv:
        gen_ff_bn_train(ref)
        return
    gen_loss(ref)
    gen_ff(ref)
    gen_ff_bn_train(ref)
    gen_groundlink(ref)
    gen_transformer(ref)
    gen_denoiser_layers(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
