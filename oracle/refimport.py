"""Import the *real* reference modules: from /root/reference in the build container, from the verbatim snapshot
``oracle/_ref`` (written by ``oracle/build_ref.py``, git-ignored, shipped by gpurun) on the GPU box, which has no
/root/reference.  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference imports ``nimblephysics`` (C++/pybind, ~=0.10.20, not installed, not vendored)
at ``src/data/AddBiomechanicsDataset.py:1`` and ``matplotlib`` at
``src/loss/RegressionLossEvaluator.py:6``.  Neither is needed by the hot path, so both are
replaced by empty stub modules before import (SURVEY §8c recipe).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IBM_REFERENCE_ROOT", "/root/reference")
SNAPSHOT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """/root/reference if present, else the oracle/_ref snapshot, else None."""
    for root in (REFERENCE_ROOT, SNAPSHOT_ROOT):
        if os.path.isfile(os.path.join(root, "src", "models", "TransformerBaseline.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def _install_stubs() -> None:
    if "nimblephysics" not in sys.modules:
        nimble = types.ModuleType("nimblephysics")
        for sub, attrs in {
            "biomechanics": ["SubjectOnDisk", "FrameList", "FramePass", "Frame", "MissingGRFReason"],
            "dynamics": ["Skeleton", "BodyNode"],
            "math": [],
        }.items():
            m = types.ModuleType(f"nimblephysics.{sub}")
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            setattr(nimble, sub, m)
            sys.modules[f"nimblephysics.{sub}"] = m
        sys.modules["nimblephysics"] = nimble
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    try:
        import wandb  # noqa: F401
    except Exception:
        w = types.ModuleType("wandb")
        w.log = lambda *a, **k: None
        sys.modules["wandb"] = w


_cache = {}


def load_reference():
    """Returns a namespace with the reference classes.  Raises RuntimeError if absent."""
    if "ns" in _cache:
        return _cache["ns"]
    root = reference_root()
    if root is None:
        raise RuntimeError(f"reference present neither at {REFERENCE_ROOT} nor as the snapshot {SNAPSHOT_ROOT} (python -m oracle.build_ref)")
    _install_stubs()
    src = os.path.join(root, "src")
    for p in (src, root):
        if p not in sys.path:
            sys.path.insert(0, p)
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        from data.AddBiomechanicsDataset import InputDataKeys, OutputDataKeys
        from models.FeedForwardRegressionBaseline import FeedForwardBaseline
        from models.Groundlink import Groundlink
        from models.TransformerBaseline import (TransformerLayer, TemporalEmbedding,
                                                SimpleAttention, TransformerBaseline)
        from loss.RegressionLossEvaluator import RegressionLossEvaluator
    ns = types.SimpleNamespace(
        InputDataKeys=InputDataKeys, OutputDataKeys=OutputDataKeys,
        FeedForwardBaseline=FeedForwardBaseline, Groundlink=Groundlink,
        TransformerLayer=TransformerLayer, TemporalEmbedding=TemporalEmbedding,
        SimpleAttention=SimpleAttention, TransformerBaseline=TransformerBaseline,
        RegressionLossEvaluator=RegressionLossEvaluator, root=root)
    _cache["ns"] = ns
    return ns
