"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/picnic_gpu.h declares.  No compute calls without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from picnic_b200 import build
    return build.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "picnic_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_error_string(lib_path):
    lib = ctypes.CDLL(lib_path)
    lib.pgpu_last_error.restype = ctypes.c_char_p
    assert lib.pgpu_abi_version() == 1
    assert isinstance(lib.pgpu_last_error(), bytes)


def test_no_cpu_fallback_without_device(lib_path):
    """Without a CUDA device pgpu_init must fail loudly; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = ctypes.CDLL(lib_path)
    lib.pgpu_last_error.restype = ctypes.c_char_p
    assert lib.pgpu_init(0) < 0
    assert b"no CPU fallback" in lib.pgpu_last_error()
    # every entry point that needs the device refuses to run
    from picnic_b200 import capi
    with pytest.raises(capi.PgpuError):
        capi.Grid(1, (8,), (0.0,), (0.25,), 2, (1,))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: no file of the product package may reference it."""
    pkg = os.path.join(ROOT, "picnic_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".H")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"\boracle\b", s) and "liboracle" in s or re.search(r"^\s*(from|import)\s+oracle", s, re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
