"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every
symbol include/ibm_b200.h declares (no compute calls without a GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from inferbiomechanics_b200 import build as b
    b.build()
    from inferbiomechanics_b200 import _lib
    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ibm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ibm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 25
    cdll = lib.load()
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in ibm_b200.h but not exported"
    assert sorted(lib.ALL_SYMBOLS) == names, "python binding table and header disagree"


def test_version_and_workspace(lib):
    cdll = lib.load()
    assert cdll.ibm_version() >= 100
    assert cdll.ibm_workspace_bytes() >= 64 * 1024


def test_no_cpu_fallback(lib):
    import torch
    from inferbiomechanics_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(lib.IbmError):
        ops.gemm(a, a, a, 8, 8, 8)


def test_sass_is_blackwell_native():
    """The GEMM object must contain tcgen05 / TMA SASS (UTCHMMA, UTMALDG, UTMASTG, LDTM), not HMMA."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    obj = os.path.join(ROOT, "inferbiomechanics_b200", "build", "gemm_sm100.o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM"):
        assert mnem in sass, mnem
    assert "HMMA.16816" not in sass
