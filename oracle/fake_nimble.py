"""A stand-in for ``nimblephysics`` that lets the REAL reference window batcher run.
TEST INFRASTRUCTURE ONLY (used by ``oracle/gen_golden.py::gen_windows`` in the build container).

``/root/reference/src/data/AddBiomechanicsDataset.py`` needs nimblephysics (~=0.10.20, C++/pybind,
not installed) only as a *reader*: ``SubjectOnDisk(path)`` with ``getNumDofs`` (105),
``getGroundForceBodies`` (107, 230), ``getNumTrials`` / ``getTrialLength`` / ``getMissingGRF``
(131-135), ``getMassKg`` (214), ``readFrames(trial, start, n, stride=…)`` (166-171) whose frames carry
``processingPasses[i].<field>`` (174-247), and the enum ``MissingGRFReason.notMissingGRF`` (133).
This module provides exactly that surface backed by the synthetic-subject dicts of
``oracle/windows.py``, so the reference class itself — not a restatement — produces ``windows`` and
the ``__getitem__`` dicts that pin ``oracle/windows.py`` and the CUDA packers.

Two processing passes per frame: pass 0 holds the subject's kinematics and contact arrays, the last
pass holds tau / residual wrench / CoM acceleration; every field of the *other* pass is the true array
plus ``DECOY`` so a reader that takes a field from the wrong pass cannot match (Dataset.py:174-175
takes inputs and contact labels from pass 0 and tau/residual/comAcc from pass -1).
"""
from __future__ import annotations

import enum
import sys
import types
from typing import Dict, List, Sequence

import numpy as np

from .windows import INPUT_ORDER, LABEL_FIRST_PASS, LABEL_LAST_PASS

DECOY = 1000.0
CANONICAL_BODIES = ("calcn_l", "calcn_r", "toes_l", "toes_r")


class MissingGRFReason(enum.Enum):
    notMissingGRF = 0
    measuredGrfZeroWhenAccelerationNonZero = 1
    unmeasuredExternalForceDetected = 2
    torqueDiscrepancy = 3
    forceDiscrepancy = 4
    notOverForcePlate = 5
    missingImpact = 6
    missingBlip = 7
    shiftGRF = 8


class FramePass:
    pass


class Frame:
    def __init__(self, passes):
        self.processingPasses = passes


_REGISTRY: Dict[str, dict] = {}


def body_names(contact_indices: Sequence[int], num_bodies: int) -> List[str]:
    """The subject's own ground-force body list such that the dataset's name lookup
    (Dataset.py:229-231) yields ``contact_indices`` against the canonical dataset order."""
    names = [f"absent_{k}" for k in range(num_bodies)]
    for b, ci in enumerate(contact_indices):
        if ci >= 0:
            names[ci] = CANONICAL_BODIES[b]
    return names


class SubjectOnDisk:
    def __init__(self, path: str):
        self._s = _REGISTRY[path]
        nb = len(self._s["contact_indices"])
        self._bodies = body_names(self._s["contact_indices"], nb)

    def getNumDofs(self) -> int:
        return int(self._s["trials"][0]["pos"].shape[1])

    def getGroundForceBodies(self) -> List[str]:
        return list(self._bodies)

    def getNumTrials(self) -> int:
        return len(self._s["trials"])

    def getTrialLength(self, trial: int) -> int:
        return len(self._s["trials"][trial]["missing"])

    def getMissingGRF(self, trial: int):
        reasons = list(MissingGRFReason)[1:]
        return [reasons[i % len(reasons)] if m else MissingGRFReason.notMissingGRF
                for i, m in enumerate(self._s["trials"][trial]["missing"])]

    def getMassKg(self) -> float:
        return float(self._s["mass"])

    def getNumProcessingPasses(self) -> int:
        return 2

    def readFrames(self, trial, startFrame, numFramesToRead, stride=1, includeSensorData=True,
                   includeProcessingPasses=True):
        tr = self._s["trials"][trial]
        L = len(tr["missing"])
        frames = []
        for k in range(numFramesToRead):
            r = startFrame + k * stride
            if r >= L:
                break
            first, last = FramePass(), FramePass()
            for f in INPUT_ORDER + LABEL_FIRST_PASS:
                setattr(first, f, tr[f][r].copy())
                setattr(last, f, tr[f][r] + DECOY)
            for f in LABEL_LAST_PASS:
                setattr(last, f, tr[f][r].copy())
                setattr(first, f, tr[f][r] + DECOY)
            frames.append(Frame([first, last]))
        return frames


def install(subjects_by_path: Dict[str, dict]) -> None:
    """Register the synthetic subjects and make ``import nimblephysics`` resolve to this fake."""
    _REGISTRY.clear()
    _REGISTRY.update(subjects_by_path)
    nimble = types.ModuleType("nimblephysics")
    bio = types.ModuleType("nimblephysics.biomechanics")
    dyn = types.ModuleType("nimblephysics.dynamics")
    mth = types.ModuleType("nimblephysics.math")
    bio.SubjectOnDisk, bio.MissingGRFReason, bio.Frame, bio.FramePass = SubjectOnDisk, MissingGRFReason, Frame, FramePass
    bio.FrameList = list
    dyn.Skeleton, dyn.BodyNode = type("Skeleton", (), {}), type("BodyNode", (), {})
    nimble.biomechanics, nimble.dynamics, nimble.math = bio, dyn, mth
    # a previously imported reference Dataset module keeps a reference to the old stub: patch it too
    for name, mod in (("nimblephysics", nimble), ("nimblephysics.biomechanics", bio),
                      ("nimblephysics.dynamics", dyn), ("nimblephysics.math", mth)):
        sys.modules[name] = mod
    ds = sys.modules.get("data.AddBiomechanicsDataset")
    if ds is not None:
        ds.nimble = nimble
