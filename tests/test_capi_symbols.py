"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/sfmgms.h declares,
and it fails loudly (no fallback) when there is no GPU.  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sfmgms.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sfmgms_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import sfm_gms_b200 as sg

    lib = sg.load_library()
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "libsfmgms.so does not export %s" % s
    assert lib.sfmgms_version() >= 100


def test_header_is_plain_c():
    """The boundary must be bindable from C / cgo / JNI: compile the header as C11."""
    import subprocess
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "sfmgms.h"\nint main(void){return SFMGMS_OK;}\n')
        subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               "-c", src, "-o", os.path.join(d, "t.o")])


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sfm_gms_b200 as sg

    with pytest.raises(sg.SfmGmsError) as e:
        sg.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under sfm_gms_b200/ or include/ may reference it."""
    bad = []
    for base in ["sfm_gms_b200", "include"]:
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"\bimport oracle\b|from oracle\b|oracle/|libsfmgms_oracle", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_host_alloc_fails_loudly_without_gpu():
    """sfmgms_host_alloc hands out page-locked memory: without a CUDA device it reports an error, never a malloc."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import sfm_gms_b200 as sg
    from sfm_gms_b200 import api

    lib = sg.load_library()
    p = ctypes.c_void_p(1)
    assert lib.sfmgms_host_alloc(1024, ctypes.byref(p)) == 5 and not p.value   # SFMGMS_ERR_CUDA
    assert lib.sfmgms_host_alloc(1024, None) == 1                                # SFMGMS_ERR_ARG
    assert lib.sfmgms_host_free(None) == 0
    with pytest.raises(sg.SfmGmsError):
        api.host_empty(16, "u1")


def test_python_mirror_constants_match_the_header():
    """api.py restates the header's option keys / kernel ids / error codes by hand: keep them equal."""
    from sfm_gms_b200 import api

    hdr = open(os.path.join(ROOT, "include", "sfmgms.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(SFMGMS_[A-Z0-9_]+)\s+(-?\d+)\b", hdr)}
    keys = [k for k in defs if k.startswith("SFMGMS_OPT_")]
    assert len(keys) >= 9 and len({defs[k] for k in keys}) == len(keys), "option keys must be distinct"
    for name in keys:
        py = name[len("SFMGMS_"):]
        assert getattr(api, py) == defs[name], name
    for name in ("SFMGMS_HAMMING_POPC", "SFMGMS_HAMMING_TC", "SFMGMS_HAMMING_FP4", "SFMGMS_HOST", "SFMGMS_DEVICE"):
        assert getattr(api, name if hasattr(api, name) else name[len("SFMGMS_"):]) == defs[name], name
