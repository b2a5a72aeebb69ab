"""The C++ drop-in classes (weiner_slamit_v2_b200/shim): they compile against mini-cv and -- in the
build container, where /root/reference exists -- against the reference's own headers, and the
reference's unchanged call sites (Frame.cc, Tracking.cc) type-check against the replacement header.
On a GPU the shim is driven exactly like Frame::ExtractORB and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from weiner_slamit_v2_b200.frames import synthetic_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "weiner_slamit_v2_b200", "shim")
PKG = os.path.join(ROOT, "weiner_slamit_v2_b200")
J = "/root/reference/oRB_SLAM2_Android/src/main/jni"
REF_INC = ["-I" + os.path.join(J, "ORB_SLAM2/include"),
           "-I/root/reference/openCVLibrary341/src/sdk/native/jni/include",
           "-I" + os.path.join(ROOT, "oracle/minicv/include/android/.."),   # android/log.h stub only
           "-I" + os.path.join(J, "Thirdparty/DBoW2/include"), "-I" + os.path.join(J, "Thirdparty/DBoW2/DLib/include"),
           "-I" + os.path.join(J, "Thirdparty/eigen3"), "-I" + os.path.join(J, "Thirdparty"), "-I" + J,
           "-I" + os.path.join(J, "ORB_SLAM2")]
needs_reference = pytest.mark.skipif(not os.path.isdir(J), reason="/root/reference not present")


def _build_driver(tmp):
    exe = os.path.join(tmp, "shim_extract")
    subprocess.check_call(["g++", "-std=c++11", "-O2", "-I" + os.path.join(ROOT, "oracle/minicv/include"),
                           "-I" + os.path.join(ROOT, "include"), "-I" + SHIM, os.path.join(SHIM, "ORBextractor.cc"),
                           os.path.join(ROOT, "tests/shim/shim_extract_main.cc"), "-o", exe, "-L" + PKG,
                           "-l:liborb_b200.so", "-Wl,-rpath," + PKG])
    return exe


def test_extractor_shim_builds_against_minicv(tmp_path, orb_lib):
    assert os.path.exists(_build_driver(str(tmp_path)))


def _stub_dir(tmp):
    d = os.path.join(tmp, "stub", "android")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "log.h"), "w") as f:
        f.write(open(os.path.join(ROOT, "oracle/minicv/include/android/log.h")).read())
    return "-I" + os.path.join(tmp, "stub")


@needs_reference
def test_shims_compile_against_reference_headers(tmp_path):
    inc = [i for i in REF_INC if "minicv" not in i] + [_stub_dir(str(tmp_path)), "-I" + os.path.join(ROOT, "include")]
    base = ["g++", "-std=c++11", "-fsyntax-only", "-w"]
    subprocess.check_call(base + inc + [os.path.join(SHIM, "ORBmatcher_b200.cc")])
    # Frame::ComputeStereoMatches reaches the extractors' device handles: built with the replacement ORBextractor.h
    subprocess.check_call(base + ["-include", os.path.join(SHIM, "ORBextractor.h")] + inc + [os.path.join(SHIM, "Frame_b200.cc")])
    subprocess.check_call(base + inc + [os.path.join(SHIM, "MapPoint_b200.cc")])
    subprocess.check_call(base + ["-I" + SHIM] + inc + [os.path.join(SHIM, "ORBextractor.cc")])


def _errors(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    return sorted(set(l.split("error:", 1)[1].strip() for l in r.stdout.splitlines() if " error:" in l))


@needs_reference
@pytest.mark.parametrize("src", ["Frame.cc", "Tracking.cc"])
def test_reference_call_sites_type_check_against_replacement_header(tmp_path, src):
    """-include puts our ORBextractor.h first; the reference's copy then vanishes behind the shared
    include guard, so Frame.cc:365 and Tracking.cc:156-162 are checked against OUR declarations.
    (Tracking.cc:285 compares a cv::Mat with nullptr, which only OpenCV 2.4.9 accepts; that one
    pre-existing diagnostic is independent of the header, so the check is: no NEW errors.)"""
    inc = [i for i in REF_INC if "minicv" not in i] + [_stub_dir(str(tmp_path))]
    base = ["g++", "-std=c++11", "-fsyntax-only", "-w"]
    target = [os.path.join(J, "ORB_SLAM2/src", src)]
    theirs = _errors(base + inc + target)
    ours = _errors(base + ["-include", os.path.join(SHIM, "ORBextractor.h")] + inc + target)
    assert ours == theirs
    if src == "Frame.cc":
        assert ours == []


@pytest.mark.gpu
def test_shim_driven_like_frame_extractorb_matches_oracle(tmp_path):
    exe = _build_driver(str(tmp_path))
    img = synthetic_frame(4)
    raw, out = os.path.join(str(tmp_path), "in.raw"), os.path.join(str(tmp_path), "out.bin")
    img.tofile(raw)
    subprocess.check_call([exe, raw, "640", "480", out])
    buf = open(out, "rb").read()
    n = int(np.frombuffer(buf, np.int32, 1)[0])
    kps = np.frombuffer(buf, O.KP_DTYPE, n, 4)
    desc = np.frombuffer(buf, np.uint8, n * 32, 4 + 28 * n).reshape(n, 32)
    orc = O.OracleExtractor()
    ko, do = orc(img)
    assert n == len(ko) and kps.tobytes() == ko.tobytes() and np.array_equal(desc, do)
    off = 4 + 60 * n
    for l in range(8):                      # mvImagePyramid with its 19-px REFLECT_101 frame
        w, h = np.frombuffer(buf, np.int32, 2, off); off += 8
        lvl = np.frombuffer(buf, np.uint8, (w + 38) * (h + 38), off).reshape(h + 38, w + 38); off += (w + 38) * (h + 38)
        assert np.array_equal(lvl, O.copy_make_border(orc.level_pixels(l), 19)), "pyramid level %d" % l
