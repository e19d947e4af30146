"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/molar_b200.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from molar_b200 import build, _capi
    build.build()
    return _capi.load()


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "molar_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from molar_b200 import _capi
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in molar_b200.h but not exported"
        assert s in _capi.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_capi.SIGNATURES) == syms


def test_no_torch_or_oracle_in_product():
    # the product path must not route through the oracle or any CPU fallback
    for dirpath, _, files in os.walk(os.path.join(ROOT, "molar_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in src and "libmolar_oracle" not in src, f
                assert "import torch" not in src, f


def test_library_has_only_sm100a_code():
    from molar_b200 import build
    log = open(os.path.join(os.path.dirname(build.LIB), "build.log")).read()
    assert "sm_100a" in log and "sm_90" not in log


def test_open_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from molar_b200 import _capi
    h = lib.mb_open(0)
    assert not h
    assert "no CPU fallback" in _capi.last_error()
    import numpy as np
    import molar_b200
    with pytest.raises(molar_b200.MolarB200Error):
        molar_b200.System(np.zeros((4, 3), np.float32))


def test_abi_version(lib):
    assert lib.mb_abi_version() == 1
