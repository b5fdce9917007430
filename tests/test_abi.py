"""The C-ABI library: loads, exports every symbol include/vrsbs.h declares, and fails loudly
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "vrsbs.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(vrsbs_[a-z_]+)\s*\(", text)))


def test_library_builds_loads_and_exports_the_header():
    import __graft_entry__ as g
    g.build()
    from vr_video_generator_b200 import _native
    lib = _native.load()
    names = _declared()
    assert names == sorted(_native.SYMBOLS), "include/vrsbs.h and _native.SYMBOLS disagree"
    for n in names:
        assert hasattr(lib, n), f"libvrsbs.so does not export {n}"
    assert lib.vrsbs_abi_version() == _native.ABI_VERSION


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import argparse

    import vr_video_generator_b200 as pkg
    from vr_video_generator_b200 import _native
    lib = _native.load()
    h = ctypes.c_void_p()
    rc = lib.vrsbs_create(ctypes.byref(h), 0, 1080, 1920, 4, 512)
    assert rc == -2 and h.value is None
    assert b"no CPU fallback" in lib.vrsbs_last_error(None)
    with pytest.raises(_native.VrsbsError):
        _native.Context(0, 1080, 1920, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.SbsProcessor(None, 0, argparse.Namespace(offset_fg=0.025, offset_bg=-0.01, offset_step_size=1))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vr-video-generator_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports oracle/"
                assert "sbs_oracle" not in src, f"{fn} references the C oracle"
