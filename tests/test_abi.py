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


def test_ctypes_structures_match_the_header(tmp_path):
    """The Python binding mirrors the header's structs by hand: sizes and the offset of every struct's last field must
    agree with what a C compiler makes of include/picnic_gpu.h (a silent mismatch would shift every later parameter)."""
    import ctypes as C
    import shutil
    import subprocess
    from picnic_b200 import capi
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    pairs = [("pgpu_grid_desc", capi.GridDesc), ("pgpu_species_desc", capi.SpeciesDesc), ("pgpu_ext_fn", capi.ExtFn),
             ("pgpu_picard_stats", capi.PicardStats), ("pgpu_halo_msg", capi.HaloMsg),
             ("pgpu_coulomb_params", capi.CoulombParams), ("pgpu_elastic_params", capi.ElasticParams)]
    src = tmp_path / "sizes.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "picnic_gpu.h"', 'int main(void) {']
    for cname, cls in pairs:
        last = cls._fields_[-1][0]
        lines.append('  printf("%s %%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));' % (cname, cname, cname, last))
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "sizes"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run([cc, "-I", inc, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    got = {l.split()[0]: (int(l.split()[1]), int(l.split()[2])) for l in out if l.strip()}
    for cname, cls in pairs:
        last = cls._fields_[-1][0]
        assert got[cname] == (C.sizeof(cls), getattr(cls, last).offset), (cname, got[cname], C.sizeof(cls))
