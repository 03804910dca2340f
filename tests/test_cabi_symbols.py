"""CPU: the C-ABI library loads without a GPU and exports exactly what include/flemingo_b200.h declares."""
import re
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (REPO / "include" / "flemingo_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fl_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    from flemingo_b200 import _cabi
    if not _cabi.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    lib = _cabi.load_library()
    declared = header_symbols()
    assert declared == sorted(_cabi.EXPORTED)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.fl_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", str(_cabi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (fl_[a-z_0-9]+)", out)))
    assert exported == declared


def test_missing_device_fails_loudly():
    """No CPU fallback: creating an engine without a CUDA device raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from flemingo_b200 import _cabi
    from flemingo_b200.engine import Engine
    with pytest.raises(_cabi.FlemingoError):
        Engine(0)


def test_product_code_never_touches_the_oracle():
    """No file under flemingo_b200/ imports, includes, links or dlopens anything from oracle/."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|placement_oracle|libplacement_oracle|oracle\.oracle|#include\s+\".*oracle|_ref/", re.M)
    for path in (REPO / "flemingo_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".h"):
            assert not pat.search(path.read_text()), path
