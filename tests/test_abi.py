"""The C-ABI library loads without a GPU and exports every symbol include/orb_b200.h declares;
argument validation and the no-device error path behave as the header says."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "orb_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(orbb200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(orb_lib):
    from weiner_slamit_v2_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(orb_lib, n), "liborb_b200.so lacks %s" % n
    assert set(names) == set(_lib.SIGNATURES), "python binding table and header disagree"


def test_keypoint_layout_is_cv_keypoint():
    from weiner_slamit_v2_b200 import KP_DTYPE
    assert KP_DTYPE.itemsize == 28
    assert [KP_DTYPE.fields[n][1] for n in KP_DTYPE.names] == [0, 4, 8, 12, 16, 20, 24]


def test_no_device_fails_loudly(orb_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from weiner_slamit_v2_b200 import ORBextractor, OrbB200Error
    assert orb_lib.orbb200_device_count() == 0
    with pytest.raises(OrbB200Error):
        ORBextractor(1000, 1.2, 8, 20, 7)


def test_argument_validation(orb_lib):
    h = C.c_void_p()
    assert orb_lib.orbb200_extractor_create(1000, 1.2, 0, 20, 7, 640, 480, 1, 0, 0, C.byref(h)) == -1      # nlevels
    assert orb_lib.orbb200_extractor_create(1000, 1.0, 8, 20, 7, 640, 480, 1, 0, 0, C.byref(h)) == -1      # scale
    assert orb_lib.orbb200_extractor_create(1000, 1.2, 8, 20, 7, 5000, 480, 1, 0, 0, C.byref(h)) == -1     # width
    assert b"invalid" in orb_lib.orbb200_last_error()
    assert orb_lib.orbb200_extractor_sync(None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "weiner_slamit_v2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".sh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "liborb_oracle" not in txt and "ref_lib" not in txt, f
