"""Dataset plumbing shared by the train / analyze commands.

The reference opens ``<dataset-home>/{train,dev}`` folders of ``.b3d`` files through nimblephysics
(/root/reference/src/cli/train.py:135-150).  That reader is outside the hot path, so here a split is a
pre-packed window store file ``<dataset-home>/<split>.ibmstore`` (WindowStore.save, written once by an export
job that does have nimblephysics), or — with ``--synthetic-windows N`` — a synthetic AddBiomechanics-shaped
store generated in HBM, which is what the tests and the benchmark use.
"""
from __future__ import annotations

import os

from ..data.window_store import WindowStore

NUM_DOFS = 23            # AddBiomechanics Rajagopal skeleton (D = 23)
NUM_JOINTS = 12


def frame_width(model_type: str, stride: int, root_history_len: int) -> int:
    """Per-frame model concat width: FeedForward uses stride*3 history columns (FeedForward…py:92-94), Groundlink and
    the denoiser root_history_len*3 (Groundlink.py:116-118)."""
    hist = stride * 3 if model_type == "feedforward" else root_history_len * 3
    return 3 * NUM_DOFS + 12 + 3 * NUM_JOINTS + 2 * hist


def open_split(args, split: str, model_type: str, device, seed: int) -> WindowStore:
    width = frame_width(model_type, args.stride, 10)
    path = os.path.join(os.path.abspath(args.dataset_home), f"{split}.ibmstore")
    n_syn = getattr(args, "synthetic_windows", 0)
    if n_syn:
        n = n_syn if split == "train" else max(n_syn // 8, 1)
        if getattr(args, "short", False):
            n = min(n, 512)
        return WindowStore.synthetic(n, args.history_len, args.stride, width, args.output_data_format, seed=seed, device=device,
                                     trial_len=max(2000, args.history_len + 64))
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found.  Export the split once with WindowStore.save(...) (needs nimblephysics to read .b3d files), "
            f"or pass --synthetic-windows N to run on synthetic AddBiomechanics-shaped windows.")
    store = WindowStore.load(path, args.history_len, args.stride, args.output_data_format, device=device)
    if store.C != width:
        raise ValueError(f"{path} holds {store.C}-column frames but model type {model_type!r} needs {width}")
    return store
