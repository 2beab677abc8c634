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
windows.npz            the reference's OWN AddBiomechanicsDataset (src/data/AddBiomechanicsDataset.py:64-139
                       index, 161-285 __getitem__) run over synthetic subjects through oracle/fake_nimble.py,
                       plus torch's DistributedSampler + DataLoader as wired at src/cli/train.py:143-150.
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
        # BASELINE configs[0] exactly: hidden [512, 512], sigmoid, batch 32 (train.py:37-53 defaults)
        "sigmoid_cfg0": dict(act="sigmoid", fmt="all_frames", hidden=[512, 512], bn=False, B=32),
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
        B = c.get("B", 6)
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


def gen_transformer_bwd(ref):
    """Autograd backward of the same composition (TransformerBaseline.py:104-148, fp64): loss = sum over the three outputs of
    <output, seeded cotangent>; every parameter gradient is frozen (fp32 copies: 208 k parameters per case)."""
    d = {}
    D = 23
    for ci, (name, B, T) in enumerate([("t20", 3, 20), ("t64", 2, 64), ("t200", 2, 200)]):
        m = ref.TransformerBaseline(D, T)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        seed = 1250 + ci
        m.load_state_dict(seeded_state_dict(shapes, seed, dtype=torch.float64))
        m.train()                                            # dropout p = 0.0: identical to eval
        x = {k: seeded_tensor((B, c, T), 2250 + ci + 10 * i, dtype=torch.float64)
             for i, (k, c) in enumerate([("pos", D), ("vel", D), ("acc", D), ("comPos", 3), ("comVel", 3), ("comAcc", 3)])}
        vecs = torch.cat([x["pos"], x["vel"], x["acc"], x["comPos"], x["comVel"], x["comAcc"]], dim=1).transpose(1, 2)
        emb = m.temporal_embedding(torch.arange(vecs.size(1))).expand(B, T, m.temporal_embedding_dim)
        vecs = torch.cat([vecs, emb], dim=2)
        for layer in m.transformer_layers:
            vecs = layer(vecs)
        output = m.fc(vecs)
        blend = m.com_attention(vecs, vecs, x["comAcc"].transpose(1, 2))
        outs = {"contact": m.contact_sigmoid(output[:, :, :2]).transpose(1, 2), "comAcc": blend.transpose(1, 2),
                "contactForces": output[:, :, 5:].transpose(1, 2)}
        loss = 0.0
        for i, (k, v) in enumerate(outs.items()):
            cot = seeded_tensor(tuple(v.shape), 2290 + ci + 10 * i, dtype=torch.float64)
            loss = loss + (v * cot).sum()
            d[f"{name}/out/{k}"] = v.detach().numpy().astype(np.float32)
        loss.backward()
        d[f"{name}/loss"] = np.array(loss.item())
        for n, p in m.named_parameters():
            d[f"{name}/grad/{n}"] = p.grad.numpy().astype(np.float32)
        d[f"{name}/meta"] = np.array([D, B, T, seed, 2250 + ci, 2290 + ci])
    np.savez_compressed(os.path.join(OUT, "transformer_bwd.npz"), **d)


def gen_ctor_variants(ref):
    """Non-default constructor arguments the reference accepts: FeedForwardBaseline(num_contact_bodies=3) (the last Linear
    grows to 45 * F outputs, the output split still takes the first 30 * F — FeedForward...py:62,116-121) and
    Groundlink(cnn_kernel, fc_depth) (Groundlink.py:20,41,51-62)."""
    d = {}
    D, T, s, B = 23, 50, 5, 5
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.FeedForwardBaseline(D, 3, T, "all_frames", "tanh", s, 10, hidden_dims=[48, 32])
    seed = 1600
    m.load_state_dict(seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed))
    m.eval()
    F = T // s
    inputs = seeded_inputs(B, F, D, s * 3, 2600)
    _, labels = seeded_out_labels(B, F, 3600)
    out = m({k: v.clone() for k, v in inputs.items()})
    ev = ref.RegressionLossEvaluator(dataset=None, split="train")
    loss = ev(None, dict(out), {k: v.clone() for k, v in labels.items()}, [], [], ns_args(*SELECTIONS["all"]))
    loss.backward()
    for k, v in out.items():
        d[f"ff_nb3/out/{k}"] = v.detach().numpy()
    d["ff_nb3/loss"] = loss.detach().numpy()
    grads_summary(m, d, "ff_nb3")
    d["ff_nb3/meta"] = np.array([D, T, s, B, seed, 2600, 3600])
    J, H = 12, 10
    for ci, (name, k, depth, fmt, B, T) in enumerate([("gl_k5_d2", 5, 2, "all_frames", 3, 30), ("gl_k3_d4", 3, 4, "last_frame", 2, 24),
                                                      ("gl_k9_d1", 9, 1, "all_frames", 2, 20)]):
        m = ref.Groundlink(D, J, H, fmt, cnn_kernel=k, fc_depth=depth)
        seed = 1610 + ci
        m.load_state_dict(seeded_state_dict({kk: tuple(v.shape) for kk, v in m.state_dict().items()}, seed))
        m.eval()
        inputs = seeded_inputs(B, T, D, H * 3, 2610 + ci)
        _, labels = seeded_out_labels(B, T if fmt == "all_frames" else 1, 3610 + ci)
        out = m({kk: v.clone() for kk, v in inputs.items()})
        ev = ref.RegressionLossEvaluator(dataset=None, split="train")
        loss = ev(None, dict(out), {kk: v.clone() for kk, v in labels.items()}, [], [], ns_args(*SELECTIONS["all"]))
        loss.backward()
        for kk, v in out.items():
            d[f"{name}/out/{kk}"] = v.detach().numpy()
        d[f"{name}/loss"] = loss.detach().numpy()
        grads_summary(m, d, name)
        d[f"{name}/meta"] = np.array([D, J, H, B, T, seed, 2610 + ci, 3610 + ci, k, depth])
    np.savez_compressed(os.path.join(OUT, "ctor_variants.npz"), **d)


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


WINDOW_CASES = {
    # name: (seed, n_subjects, T, stride, format, hist_cols, max_len, n_sampled_windows)
    "ff_t50s5_all": (11, 6, 50, 5, "all_frames", 15, 170, 10),
    "gl_t50s1_last": (12, 3, 50, 1, "last_frame", 30, 130, 4),
    "t20s5_last": (13, 5, 20, 5, "last_frame", 15, 90, 8),
    "t7s3_all": (14, 4, 7, 3, "all_frames", 9, 40, 8),       # T % stride != 0: 3 frames tested, 2 read
}
WINDOW_LABEL_KEYS = ("tau", "residualWrenchInRootFrame", "comAccInRootFrame", "groundContactWrenchesInRootFrame",
                     "groundContactCenterOfPressureInRootFrame", "groundContactTorqueInRootFrame",
                     "groundContactForceInRootFrame")


def window_digest(items) -> str:
    """sha256 over every window's input arrays (model concat order) then label arrays (WINDOW_LABEL_KEYS order)."""
    import hashlib
    from .windows import INPUT_ORDER
    h = hashlib.sha256()
    for inputs, labels in items:
        for k in INPUT_ORDER:
            h.update(np.ascontiguousarray(np.asarray(inputs[k], dtype=np.float32)).tobytes())
        for k in WINDOW_LABEL_KEYS:
            h.update(np.ascontiguousarray(np.asarray(labels[k], dtype=np.float32)).tobytes())
    return h.hexdigest()


def reference_dataset(subjects, T, s, fmt, tmp):
    """The reference's own AddBiomechanicsDataset (Dataset.py:64-139) over ``subjects`` through the fake nimblephysics reader:
    one empty ``.b3d`` per subject under ``tmp`` (plus a file the class must skip: "vander", Dataset.py:89, and a non-.b3d)."""
    import contextlib
    import io
    from . import fake_nimble
    load_reference()                                   # puts /root/reference/src on sys.path
    n_subj = len(subjects)
    os.makedirs(os.path.join(tmp, "grp_a"))
    os.makedirs(os.path.join(tmp, "grp_b"))
    for i in range(n_subj):
        open(os.path.join(tmp, "grp_a" if i % 2 else "grp_b", f"subj{i:02d}.b3d"), "w").close()
    open(os.path.join(tmp, "grp_a", "VanDerZee2022_x.b3d"), "w").close()     # skipped: "vander" (Dataset.py:89)
    open(os.path.join(tmp, "grp_b", "notes.txt"), "w").close()               # skipped: not .b3d
    found = [os.path.join(r, f) for r, _, fs in os.walk(tmp) for f in fs
             if f.endswith(".b3d") and "vander" not in f.lower()]
    assert len(found) == n_subj
    # subject k of the synthetic list is the k-th file the reference's own os.walk discovers
    fake_nimble.install({p: subjects[k] for k, p in enumerate(found)})
    from data.AddBiomechanicsDataset import AddBiomechanicsDataset
    with contextlib.redirect_stdout(io.StringIO()):
        ds = AddBiomechanicsDataset(tmp, T, geometry_folder="", stride=s, output_data_format=fmt,
                                    skip_loading_skeletons=True)
    assert ds.subject_paths == found
    return ds


def gen_windows(ref_unused=None):
    """Run the reference's AddBiomechanicsDataset itself over synthetic subjects (fake nimblephysics reader)."""
    import tempfile
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    from .windows import make_synthetic_subjects
    d = {}
    for name, (seed, n_subj, T, s, fmt, hist, max_len, n_samp) in WINDOW_CASES.items():
        subjects = make_synthetic_subjects(seed, n_subj, T, hist_cols=hist, max_len=max_len)
        with tempfile.TemporaryDirectory() as tmp:
            ds = reference_dataset(subjects, T, s, fmt, tmp)
            N = len(ds)
            d[f"{name}/windows"] = np.asarray(ds.windows, dtype=np.int32).reshape(N, 3)
            d[f"{name}/meta"] = np.array([seed, n_subj, T, s, hist, max_len, ds.num_dofs, ds.num_contact_bodies])
            d[f"{name}/format"] = np.array(fmt)
            d[f"{name}/contact_bodies"] = np.array(ds.contact_bodies)
            items = [ds[i] for i in range(N)]
            d[f"{name}/digest"] = np.array(window_digest((it[0], it[1]) for it in items))
            pick = np.unique(np.linspace(0, N - 1, n_samp).astype(np.int64))
            d[f"{name}/sample_idx"] = pick
            for i in pick:
                inp, lab, si, ti = items[int(i)]
                assert (si, ti) == tuple(ds.windows[int(i)][:2])
                for k, v in inp.items():
                    d[f"{name}/w{int(i)}/in/{k}"] = v.numpy()
                for k, v in lab.items():
                    d[f"{name}/w{int(i)}/label/{k}"] = v.numpy()
            # train.py:143-150: DistributedSampler(shuffle=False, drop_last=True) + DataLoader(batch_size), rank 1 of 3
            sampler = DistributedSampler(ds, num_replicas=3, rank=1, shuffle=False, drop_last=True)
            d[f"{name}/sampler_r1w3"] = np.asarray(list(sampler), dtype=np.int64)
            bs = 5
            subj_idx, trial_idx, sizes, pos_sum = [], [], [], []
            for batch in DataLoader(ds, batch_size=bs, sampler=sampler, num_workers=0):
                inputs, labels, bsi, bti = batch
                subj_idx.append(bsi.numpy()); trial_idx.append(bti.numpy()); sizes.append(len(bsi))
                pos_sum.append(inputs["pos"].double().sum().item())
            d[f"{name}/loader_bs5_sizes"] = np.asarray(sizes, dtype=np.int64)
            d[f"{name}/loader_bs5_subject"] = np.concatenate(subj_idx) if subj_idx else np.zeros(0, np.int64)
            d[f"{name}/loader_bs5_trial"] = np.concatenate(trial_idx) if trial_idx else np.zeros(0, np.int64)
            d[f"{name}/loader_bs5_pos_sum"] = np.asarray(pos_sum)
    np.savez_compressed(os.path.join(OUT, "windows.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)      # deterministic reduction order for the frozen vectors
    ref = load_reference()
    if "--only-bn-train" in __import__("sys").argv:
        gen_ff_bn_train(ref)
        return
    if "--only-ff" in __import__("sys").argv:
        gen_ff(ref)
        return
    if "--only-ctor-variants" in __import__("sys").argv:
        gen_ctor_variants(ref)
        return
    if "--only-transformer-bwd" in __import__("sys").argv:
        gen_transformer_bwd(ref)
        return
    if "--only-windows" in __import__("sys").argv:
        gen_windows(ref)
        return
    gen_loss(ref)
    gen_ff(ref)
    gen_ff_bn_train(ref)
    gen_groundlink(ref)
    gen_transformer(ref)
    gen_transformer_bwd(ref)
    gen_ctor_variants(ref)
    gen_denoiser_layers(ref)
    gen_windows(ref)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
