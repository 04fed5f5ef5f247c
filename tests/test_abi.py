"""The C-ABI library builds, loads without a GPU and exports every symbol include/orbx.h declares.  CPU only
(no compute call is made here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_extractor_surface():
    names = declared_functions()
    for need in ("orbx_extractor_create", "orbx_extractor_run_host", "orbx_extractor_run_device", "orbx_extractor_pyramid",
                 "orbx_extractor_tables", "orbx_extractor_destroy"):
        assert need in names


def test_library_exports_every_declared_symbol():
    from orbx import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(L, n)]
    assert not missing, missing
    assert L.orbx_version() >= 100


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from orbx import _lib
    from orbx.extractor import ORBextractor
    with pytest.raises(_lib.OrbxError) as ei:
        ORBextractor(1000, 1.2, 8, 20, 7)
    assert ei.value.status == -3


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "active-orb-slam2_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(d, f), errors="replace").read()
                assert "oracle_py" not in txt and "orbx_oracle" not in txt and "liborbx_oracle" not in txt, os.path.join(d, f)
