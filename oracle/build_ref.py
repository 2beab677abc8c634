"""Recipe for ``oracle/_ref``: the reference's own hot-path modules, copied VERBATIM from where they lie under
/root/reference into ``oracle/_ref/`` (git-ignored, so no reference source enters the history; NOT gpurun-ignored, so the
snapshot travels to the GPU box, which has no /root/reference).  TEST / BASELINE INFRASTRUCTURE ONLY.

    python -m oracle.build_ref        # also run by __graft_entry__.build() when /root/reference is present

The reference is pure Python (no build system, nothing to compile): "building" it is selecting the five files the path
needs.  ``oracle/refimport.load_reference()`` imports them (with empty ``nimblephysics`` / ``matplotlib`` stubs)
from /root/reference when that exists and from ``oracle/_ref`` otherwise; bench.py's CPU arms then time the REAL
``FeedForwardBaseline`` / ``Groundlink`` / ``TransformerLayer`` / ``RegressionLossEvaluator`` on the box's host cores.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("IBM_REFERENCE_ROOT", "/root/reference")
FILES = [
    "src/__init__.py",
    "src/data/__init__.py", "src/data/AddBiomechanicsDataset.py",
    "src/models/__init__.py", "src/models/FeedForwardRegressionBaseline.py", "src/models/Groundlink.py",
    "src/models/TransformerBaseline.py",
    "src/loss/__init__.py", "src/loss/RegressionLossEvaluator.py",
]


def build(verbose: bool = False) -> str:
    if not os.path.isdir(os.path.join(SOURCE, "src", "models")):
        raise RuntimeError(f"reference not present at {SOURCE}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
            manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
            manifest[rel] = "(empty package marker)"
        else:
            raise RuntimeError(f"reference file missing: {src}")
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SOURCE, "files": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} files from {SOURCE}")
    return DEST


if __name__ == "__main__":
    build(verbose=True)
