"""The C-ABI library loads and exports every symbol include/mamdr_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mamdr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mamdr_[a-z0-9_]+)\s*\(", text)))


def test_build_and_exports():
    from mamdr_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libmamdr_b200.so does not export %s" % name
    # the ctypes binding covers exactly the declared surface
    assert sorted(_lib.SIGNATURES) == declared
    assert _lib.load().mamdr_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header():
    from mamdr_b200 import _lib
    # mamdr_mlp_desc: 4 + 12 + 32 + 4 + 4 (+0 pad) + 16 + 4*4 + 3*8 + 64 + 64 + 16 + 8
    assert ctypes.sizeof(_lib.MlpDesc) == 264
    assert ctypes.sizeof(_lib.Batch) == 56
    assert _lib.load().mamdr_opt_state_bytes() == 32


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mamdr_b200 import _lib
    with pytest.raises(_lib.MamdrError):
        _lib.Context(0)
    from mamdr_b200.engine import MLPModel
    with pytest.raises(RuntimeError):
        MLPModel(8, 8, 2, device="cpu", user_table=None, item_table=None)


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mamdr_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(f)
    src = open(os.path.join(ROOT, "run.py")).read()
    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M)
    assert not bad, bad
