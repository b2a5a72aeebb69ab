"""Pins the oracle's OpenCV primitive restatements bit-for-bit against cv2 (4.13 in this image):
resize INTER_LINEAR, FAST-9/16 + NMS, GaussianBlur 7x7 sigma 2, copyMakeBorder REFLECT_101,
fastAtan2; and its sinf/cosf restatement against libm over every float in [0, 2*pi]."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as O
from weiner_slamit_v2_b200.frames import low_contrast_frame, synthetic_frame

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


def _kp_array(kps):
    return np.array([(k.pt[0], k.pt[1], k.response) for k in kps], np.float32).reshape(-1, 3)


def test_resize_chain_matches_cv2():
    inv = O.OracleExtractor().inv_scale_factors
    for (w, h) in ((640, 480), (1280, 720), (752, 480)):
        prev = synthetic_frame(1, w, h)
        for l in range(1, 8):
            dw, dh = O.lib().orc_round(np.float32(w) * inv[l]), O.lib().orc_round(np.float32(h) * inv[l])
            ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(ref, O.resize_linear(prev, dw, dh)), (w, h, l)
            prev = ref


def test_resize_random_sizes_matches_cv2():
    rng = np.random.default_rng(0)
    for _ in range(25):
        sw, sh = rng.integers(40, 400, 2)
        dw, dh = int(sw / rng.uniform(1.05, 1.9)), int(sh / rng.uniform(1.05, 1.9))
        src = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        assert np.array_equal(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), O.resize_linear(src, dw, dh))


@pytest.mark.parametrize("th", [20, 7, 0, 40])
def test_fast_matches_cv2_order_and_response(th):
    rng = np.random.default_rng(1)
    imgs = [synthetic_frame(2), low_contrast_frame(0), rng.integers(0, 256, (120, 160)).astype(np.uint8),
            rng.integers(90, 140, (97, 131)).astype(np.uint8)]
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for img in imgs:
        mine = O.fast(img, th, True)
        m = np.stack([mine["x"], mine["y"], mine["response"]], 1) if len(mine) else np.zeros((0, 3), np.float32)
        assert np.array_equal(_kp_array(det.detect(img)), m)


def test_fast_on_cell_sized_windows_matches_cv2():
    rng = np.random.default_rng(2)
    img = synthetic_frame(3)
    for _ in range(150):
        x, y = rng.integers(0, 600), rng.integers(0, 440)
        w, h = rng.integers(4, 40, 2)
        win = np.ascontiguousarray(img[y:y + h, x:x + w])
        for th in (20, 7):
            det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True)
            mine = O.fast(win, th, True)
            m = np.stack([mine["x"], mine["y"], mine["response"]], 1) if len(mine) else np.zeros((0, 3), np.float32)
            assert np.array_equal(_kp_array(det.detect(win)), m)


def test_gaussian_blur_matches_cv2():
    rng = np.random.default_rng(3)
    for img in (synthetic_frame(2), rng.integers(0, 256, (64, 80)).astype(np.uint8),
                rng.integers(0, 256, (134, 179)).astype(np.uint8), rng.integers(0, 256, (8, 9)).astype(np.uint8)):
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, O.gaussian_blur7(img, 0))


def test_gaussian_blur_249_variant_differs_only_slightly():
    img = synthetic_frame(2)
    a, b = O.gaussian_blur7(img, 0).astype(int), O.gaussian_blur7(img, 1).astype(int)
    assert np.abs(a - b).max() <= 3 and (a != b).any()      # taps sum to 257 instead of 256: ~+0.8% brighter


def test_copy_make_border_matches_cv2():
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (30, 41)).astype(np.uint8)
    assert np.array_equal(cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101), O.copy_make_border(img, 19))


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(5)
    ys, xs = rng.integers(-3000000, 3000000, 20000), rng.integers(-3000000, 3000000, 20000)
    ys[:100] = 0; xs[100:200] = 0; xs[:10] = 0
    for y, x in zip(ys, xs):
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(O.fast_atan2(y, x))


def test_sincosf_restatement_matches_libm_exhaustively():
    """Every float in [0, 2*pi] (1.09e9 values): the double-polynomial restatement == glibc sinf/cosf."""
    src = r'''
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
void orc_sincosf(float, float*, float*);
int main(void){ float lim=6.2832f; uint32_t ul; memcpy(&ul,&lim,4); long bad=0;
#pragma omp parallel for reduction(+:bad) schedule(static)
 for(uint32_t u=0;u<=ul;u++){ float y; memcpy(&y,&u,4); float s,c; orc_sincosf(y,&s,&c); float rs=sinf(y), rc=cosf(y);
   if(memcmp(&s,&rs,4)||memcmp(&c,&rc,4)) bad++; }
 printf("%ld\n",bad); return 0; }
'''
    O.build_oracle()
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", p, "-o", exe,
                               "-L" + O.ORACLE_DIR, "-l:liborb_oracle.so", "-Wl,-rpath," + O.ORACLE_DIR, "-lm"])
        assert int(subprocess.check_output([exe]).strip()) == 0


@pytest.mark.parametrize("dist", [[0.262383, -0.953104, -0.005358, 0.002628, 1.163314],     # the reference's camera (Tracking.cc:101-111)
                                  [0.262383, -0.953104, -0.005358, 0.002628, 0.0], [-0.28, 0.07, 2e-4, 2e-5, 0.0]])
def test_undistort_points_matches_cv2(dist):
    K = np.array([526.69, 540.36, 313.07, 238.39], np.float32)
    d = np.array(dist, np.float32)
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-50, 700, 20000), rng.uniform(-50, 530, 20000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    dc = d[:4].reshape(4, 1) if d[4] == 0 else d.reshape(5, 1)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), Km, dc, None, Km).reshape(-1, 2)
    assert np.array_equal(ref, O.undistort_points(pts, K, d))
    b = O.image_bounds(640, 480, K, d)
    assert b[0] == min(ref[0, 0], ref[2, 0]) and b[2] == max(ref[1, 0], ref[3, 0])
    assert b[1] == min(ref[0, 1], ref[1, 1]) and b[3] == max(ref[2, 1], ref[3, 1])


def test_small_float_gemm_matches_cv2():
    """x3Dc = Rcw*x3Dw + tcw (S/ORBmatcher.cc:1363): cv::gemm's small-matrix path = float products and sums in
    source order; the oracle, oracle/cv341_stubs.cc and the CUDA kernel all use exactly that."""
    rng = np.random.default_rng(0)
    f = np.float32
    for _ in range(5000):
        Rm = rng.normal(0, 1, (3, 3)).astype(np.float32); x = rng.normal(0, 5, (3, 1)).astype(np.float32)
        t = rng.normal(0, 2, (3, 1)).astype(np.float32)
        ref = cv2.gemm(Rm, x, 1.0, t, 1.0)
        mine = np.array([[f(f(f(Rm[i, 0] * x[0, 0]) + f(Rm[i, 1] * x[1, 0])) + f(Rm[i, 2] * x[2, 0])) + t[i, 0]] for i in range(3)], np.float32)
        assert np.array_equal(ref, mine)


def test_logf_restatement_is_bit_identical_to_libm_on_a_sample():
    """orc_logf (glibc's algorithm around the table in include/orb_b200_logf.inc) against this machine's libm:
    all 2,139,095,039 positive finite floats were checked once with a C loop; here a stratified
    sample (every exponent, random mantissas) plus the special values."""
    import ctypes as C
    L = O.lib()
    L.orc_logf.argtypes = [C.c_float]; L.orc_logf.restype = C.c_float
    libm = C.CDLL("libm.so.6")
    libm.logf.argtypes = [C.c_float]; libm.logf.restype = C.c_float
    rng = np.random.default_rng(5)
    bits = (np.repeat(np.arange(0, 255, dtype=np.uint32), 120) << 23) | rng.integers(0, 1 << 23, 255 * 120, dtype=np.uint32)
    bits = np.concatenate([bits, np.array([1, 0x007fffff, 0x00800000, 0x3f7fffff, 0x3f800000, 0x3f800001, 0x7f7fffff], np.uint32)])
    xs = bits.view(np.float32)
    for x in xs:
        a, b = np.float32(L.orc_logf(float(x))), np.float32(libm.logf(float(x)))
        assert a.tobytes() == b.tobytes(), (x, a, b)
    # PredictScale on exact powers of the scale factor sits on the ceil boundary: same level as the C library gives
    L.orc_predict_scale.argtypes = [C.c_float] * 3; L.orc_predict_scale.restype = C.c_int
    ls = np.float32(libm.logf(np.float32(1.2)))
    for k in range(8):
        mx = np.float32(np.float32(1.2) ** k)
        want = int(np.ceil(np.float32(np.float32(libm.logf(float(np.float32(mx / np.float32(1.0))))) / ls)))
        assert L.orc_predict_scale(float(mx), 1.0, float(ls)) == want
