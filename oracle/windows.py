"""CPU restatement of the reference window batcher.  TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/src/data/AddBiomechanicsDataset.py``:

* ``enumerate_windows``   ← window index construction in ``__init__`` (121-139)
* ``sampler_indices``     ← ``DistributedSampler(shuffle=False, drop_last=True)`` as used at
                            ``src/cli/train.py:143`` + ``DataLoader(batch_size)`` (train.py:144)
* ``get_window``          ← ``__getitem__`` (161-285)
* ``pack_inputs``         ← the concat order of ``FeedForwardBaseline.forward``
                            (``src/models/FeedForwardRegressionBaseline.py:97-108``) and
                            ``Groundlink.forward`` (``src/models/Groundlink.py:122-133``)

Pinning: the reference class ITSELF runs in the build container over these synthetic subjects through
``oracle/fake_nimble.py`` (a reader-only stand-in for nimblephysics); ``oracle/gen_golden.py::gen_windows``
freezes its ``windows`` list, ``__getitem__`` dicts (sampled windows verbatim, all windows by digest) and
torch's DistributedSampler/DataLoader order in ``tests/golden/windows.npz``.  ``tests/test_windows_golden.py``
holds this restatement (CPU) and the CUDA index/packers (``-m gpu``) to that fixture bit for bit;
``tests/test_oracle_windows.py`` adds hand-computed cases.

A synthetic "subject" stands in for ``nimble.biomechanics.SubjectOnDisk``: a dict with
``mass`` (float), ``contact_indices`` (for each dataset contact body, its index among the
subject's ground-force bodies or -1; Dataset.py:229-231), and ``trials``: a list of dicts with
``missing`` (bool[L]) and per-frame float arrays named by FRAME_FIELDS.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

# Input keys in MODEL concat order (FeedForward…py:97-108); value = per-frame width given (D, hist)
INPUT_ORDER = (
    "pos", "vel", "acc",
    "rootLinearVelInRootFrame", "rootAngularVelInRootFrame",
    "rootLinearAccInRootFrame", "rootAngularAccInRootFrame",
    "jointCentersInRootFrame",
    "rootPosHistoryInRootFrame", "rootEulerHistoryInRootFrame",
)
LABEL_LAST_PASS = ("tau", "residualWrenchInRootFrame", "comAccInRootFrame")
LABEL_FIRST_PASS = ("groundContactWrenchesInRootFrame", "groundContactCenterOfPressureInRootFrame",
                    "groundContactTorqueInRootFrame", "groundContactForceInRootFrame")


def input_widths(num_dofs: int, hist_cols: int, num_joints: int = 12) -> Dict[str, int]:
    return {
        "pos": num_dofs, "vel": num_dofs, "acc": num_dofs,
        "rootLinearVelInRootFrame": 3, "rootAngularVelInRootFrame": 3,
        "rootLinearAccInRootFrame": 3, "rootAngularAccInRootFrame": 3,
        "jointCentersInRootFrame": 3 * num_joints,
        "rootPosHistoryInRootFrame": hist_cols, "rootEulerHistoryInRootFrame": hist_cols,
    }


def make_synthetic_subjects(seed: int, num_subjects: int, window_size: int, num_dofs: int = 23,
                            num_contact_bodies: int = 2, hist_cols: int = 30,
                            max_len: int = 400, trials_per_subject: Tuple[int, int] = (1, 4),
                            missing_start_p: float = 0.02) -> List[dict]:
    """SURVEY §8d synthetic subjects: trial lengths U{T+2 … max_len}, missing-GRF runs with
    Bernoulli(p) starts and length U{1…50}; some trials deliberately shorter than the window."""
    rng = np.random.default_rng(seed)
    subjects = []
    w = input_widths(num_dofs, hist_cols)
    for s in range(num_subjects):
        trials = []
        for t in range(int(rng.integers(trials_per_subject[0], trials_per_subject[1] + 1))):
            if rng.random() < 0.15:
                L = int(rng.integers(1, window_size + 2))      # too short → contributes nothing
            else:
                L = int(rng.integers(window_size + 2, max_len + 1))
            missing = np.zeros(L, dtype=bool)
            i = 0
            while i < L:
                if rng.random() < missing_start_p:
                    run = int(rng.integers(1, 51))
                    missing[i:i + run] = True
                    i += run
                else:
                    i += 1
            tr = {"missing": missing}
            for k, c in w.items():
                tr[k] = rng.standard_normal((L, c))
            tr["tau"] = rng.standard_normal((L, num_dofs))
            tr["residualWrenchInRootFrame"] = rng.standard_normal((L, 6))
            tr["comAccInRootFrame"] = rng.standard_normal((L, 3))
            nb = num_contact_bodies
            tr["groundContactWrenchesInRootFrame"] = rng.standard_normal((L, 6 * nb)) * 300.0
            tr["groundContactCenterOfPressureInRootFrame"] = rng.standard_normal((L, 3 * nb))
            tr["groundContactTorqueInRootFrame"] = rng.standard_normal((L, 3 * nb)) * 30.0
            tr["groundContactForceInRootFrame"] = rng.standard_normal((L, 3 * nb)) * 700.0
            trials.append(tr)
        order = list(range(num_contact_bodies))
        if s % 3 == 1:
            order = order[::-1]             # subject lists its force bodies right-then-left
        if s % 5 == 4:
            order[-1] = -1                  # one contact body absent for this subject
        subjects.append({"mass": float(rng.uniform(45.0, 110.0)), "contact_indices": order,
                         "trials": trials})
    return subjects


def enumerate_windows(subjects: Sequence[dict], window_size: int, stride: int) -> List[Tuple[int, int, int]]:
    """Dataset.py:131-139.  Order: subject → trial → window_start."""
    windows: List[Tuple[int, int, int]] = []
    for i, subj in enumerate(subjects):
        for j, tr in enumerate(subj["trials"]):
            missing = tr["missing"]
            L = len(missing)
            for ws in range(max(L - window_size - 1, 0)):
                if not any(missing[ws:ws + window_size:stride]):
                    assert ws + window_size < L
                    windows.append((i, j, ws))
    return windows


def sampler_indices(n: int, world_size: int, rank: int) -> List[int]:
    """DistributedSampler(shuffle=False, drop_last=True): rank r gets r::W of the first
    floor(N/W)*W indices (torch/utils/data/distributed.py; SURVEY §8e [probed])."""
    total = (n // world_size) * world_size
    return list(range(total))[rank:total:world_size]


def batches(indices: Sequence[int], batch_size: int) -> List[List[int]]:
    """DataLoader default drop_last=False: the last partial batch is kept (train.py:144)."""
    return [list(indices[i:i + batch_size]) for i in range(0, len(indices), batch_size)]


def get_window(subjects: Sequence[dict], window: Tuple[int, int, int], window_size: int, stride: int,
               output_data_format: str, num_contact_bodies: int):
    """Dataset.py:161-285.  Returns (inputs, labels) dicts of float32 arrays (F,C)/(F',C)."""
    si, ti, ws = window
    subj = subjects[si]
    tr = subj["trials"][ti]
    F = window_size // stride
    rows = ws + stride * np.arange(F)                      # readFrames(trial, ws, F, stride)
    inputs = {k: tr[k][rows].astype(np.float32) for k in INPUT_ORDER}
    lab_rows = rows if output_data_format == "all_frames" else rows[-1:]
    labels = {k: tr[k][lab_rows].astype(np.float32) for k in LABEL_LAST_PASS}
    mass = np.float32(subj["mass"])
    widths = {"groundContactWrenchesInRootFrame": 6, "groundContactCenterOfPressureInRootFrame": 3,
              "groundContactTorqueInRootFrame": 3, "groundContactForceInRootFrame": 3}
    for k in LABEL_FIRST_PASS:
        wd = widths[k]
        out = np.zeros((len(lab_rows), wd * num_contact_bodies), dtype=np.float32)
        src = tr[k][lab_rows].astype(np.float32)
        for b in range(num_contact_bodies):
            ci = subj["contact_indices"][b]
            if ci >= 0:
                blk = src[:, wd * ci: wd * ci + wd]
                if k != "groundContactCenterOfPressureInRootFrame":
                    blk = blk / mass                     # fp32 / fp32 (Dataset.py:248-261)
                out[:, wd * b: wd * b + wd] = blk
        labels[k] = out
    return inputs, labels


def pack_inputs(inputs: Dict[str, np.ndarray], flatten: bool) -> np.ndarray:
    """Concat in model order; (F, C_frame) for Groundlink, (F*C_frame,) frame-major for FF."""
    x = np.concatenate([inputs[k] for k in INPUT_ORDER], axis=-1)
    return x.reshape(-1) if flatten else x


def pack_labels30(labels: Dict[str, np.ndarray]) -> np.ndarray:
    """The 30-channel per-frame label row [CoP6 | F6 | tau6 | W12] used by the packed path
    (same channel order as Groundlink's output slices, Groundlink.py:151-156)."""
    return np.concatenate([labels["groundContactCenterOfPressureInRootFrame"],
                           labels["groundContactForceInRootFrame"],
                           labels["groundContactTorqueInRootFrame"],
                           labels["groundContactWrenchesInRootFrame"]], axis=-1)
