"""CPU: libavdn.so builds, loads and exports every symbol include/avdn.h declares.
No compute entry point is called here (no GPU in the CPU tier)."""
import os
import re

from avdn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "avdn.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(avdn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    h = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 7
    for sym in declared:
        assert hasattr(h, sym), f"{sym} declared in avdn.h but not exported"
    # and the Python binding table covers the header
    assert set(declared) == set(_lib.exported_symbols())


def test_abi_version_and_error_string(built_lib):
    h = _lib.lib()
    assert h.avdn_abi_version() >= 1
    assert isinstance(_lib.last_error(), str)


def test_bad_arguments_are_rejected_without_a_gpu(built_lib):
    h = _lib.lib()
    # null pointers are refused before any CUDA call is made
    assert h.avdn_pack_tile(None, None, 0, 8, 8, None, None) == -1
    assert "null" in _lib.last_error()
    assert h.avdn_render_views(None, 0, None, None, 4, None, None, None, None, None, None) == -1
    assert h.avdn_homography_from_corners(None, 0, None, None) == 0      # empty batch is a no-op


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "aerial-vision-and-dialog-navigation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
