"""The drop-in C++ classes EXECUTED, not just compiled: oracle/_ref/libshim_matcher.so (oracle/Makefile, `make shimm`)
links weiner_slamit_v2_b200/shim/{ORBmatcher_b200,Frame_b200,MapPoint_b200,ORBextractor}.cc into the reference's own
ORBmatcher.cc / Frame.cc / MapPoint.cc / KeyFrame.cc objects (every method the shim defines replaces the reference's
body, the rest stays the reference's) and is driven by the SAME harness that drives libref_matcher.so.  So
`ORBmatcher(0.9, true).SearchForInitialization(F1, F2, prev, m12, 100)` etc. run here on real ORB_SLAM2::Frame /
MapPoint / KeyFrame objects (interface I/ORBmatcher.h:41-83; callers S/Tracking.cc:799-800, 1115-1121, 1451-1462),
flatten them, search on the GPU through liborb_b200.so and write back into F.mvpMapPoints / vnMatches12 /
vbPrevMatched -- and are compared with the vectors the reference's own bodies produced (tests/golden/ref_match_*.npz,
tools/gen_golden.py) and with the oracle."""
import ctypes as C
import hashlib
import os
import threading

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R
from weiner_slamit_v2_b200.frames import stereo_right_frame, synthetic_frame
from weiner_slamit_v2_b200.workloads import (SCALE_FACTORS_8, bow_pair, fuse_frame, init_pair, motion_frame,
                                             observed_descriptors, projection_frame, relocalisation_frame, sim3_pair,
                                             synthetic_vocabulary, triangulation_pair, vocabulary_features,
                                             write_vocabulary_text)

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.shim_available(), reason="oracle/_ref/libshim_matcher.so not built")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BOUNDS = (-13.7, -9.2, 661.3, 492.8)
_u8p, _f32p, _i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def gold(name):
    return np.load(os.path.join(GOLD, name))


def test_descriptor_distance_shim():
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    with R.shim_bodies():
        for i in range(0, 64, 2):
            assert R.ref_descriptor_distance(d[i], d[i + 1]) == O.descriptor_distance(d[i], d[i + 1])


def test_search_for_initialization_shim_reproduces_reference():
    g = gold("ref_match_init.npz")
    with R.shim_bodies():
        for i in range(int(g["count"])):
            idx, n, brute, window = (int(v) for v in g["cfg_%d" % i])
            p = init_pair(idx, n=n, brute_force=bool(brute))
            cnt, m12, prev = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, window)
            assert cnt == int(g["n_%d" % i]) and np.array_equal(m12, g["m12_%d" % i]) and np.array_equal(prev, g["prev_%d" % i]), i
        # a second round on the updated vbPrevMatched, an empty side, and no orientation check: against the oracle
        p = init_pair(410, n=700)
        cnt, m12, prev = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, 100)
        on, om, op = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, 100)
        assert cnt == on and np.array_equal(m12, om) and np.array_equal(prev, op)
        cnt2, m2, prev2 = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], prev, (0, 0, 640, 480), 0.9, False, 40)
        on2, om2, op2 = O.search_for_initialization(p[0], p[1], p[2], p[3], op, (0, 0, 640, 480), 0.9, False, 40)
        assert cnt2 == on2 and np.array_equal(m2, om2) and np.array_equal(prev2, op2)
        cnt0, m0, _ = R.ref_search_for_initialization(p[0][:0], p[1][:0], p[2], p[3], p[4][:0], (0, 0, 640, 480), 0.9, True, 100)
        assert cnt0 == 0 and len(m0) == 0


def test_search_by_projection_shim_reproduces_reference():
    g = gold("ref_match_proj.npz")
    with R.shim_bodies():
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            kp, kd, mp = projection_frame(int(c[0]), int(c[1]), int(c[2]))
            cnt, kpmp = R.ref_search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, float(c[3]))
            assert cnt == int(g["n_%d" % i]) and np.array_equal(kpmp, g["kpmp_%d" % i]), i
        # keypoints that already hold a map point (of this list, and foreign ones with and without observations)
        kp, kd, mp = projection_frame(510, 1500, 5000)
        rng = np.random.default_rng(9)
        pre = np.full(len(kp), -1, np.int32); obs = np.zeros(len(kp), np.int32)
        own = rng.choice(len(kp), 100, replace=False); pre[own] = rng.integers(0, 5000, 100)
        foreign = rng.choice(np.setdiff1d(np.arange(len(kp)), own), 100, replace=False); pre[foreign] = -2; obs[foreign] = rng.integers(0, 3, 100)
        cnt, kpmp = R.ref_search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, 3.0, pre, obs)
        oc, okp = O.search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, 3.0, pre, obs)
        assert cnt == oc and np.array_equal(kpmp, okp)


def test_search_by_projection_last_frame_and_keyframe_shim_reproduce_reference():
    g = gold("ref_match_lastframe.npz")
    with R.shim_bodies():
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            cnt, kpmp = R.ref_search_by_projection_last_frame(motion_frame(int(c[0])), SCALE_FACTORS_8, BOUNDS, float(c[1]), bool(c[2]))
            assert cnt == int(g["n_%d" % i]) and np.array_equal(kpmp, g["kpmp_%d" % i]), i
        g = gold("ref_match_keyframe.npz")
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            cnt, kpmp, _ = R.ref_search_by_projection_keyframe(relocalisation_frame(int(c[0])), SCALE_FACTORS_8, BOUNDS, float(c[1]), int(c[2]), bool(c[3]))
            assert cnt == int(g["n_%d" % i]) and np.array_equal(kpmp, g["kpmp_%d" % i]), i


def test_search_by_bow_and_triangulation_shim_reproduce_reference():
    with R.shim_bodies():
        g = gold("ref_match_bow.npz")
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            cnt, m = R.ref_search_by_bow(bow_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3])), float(c[4]), bool(c[5]))
            assert cnt == int(g["n_%d" % i]) and np.array_equal(m, g["m_%d" % i]), i
        g = gold("ref_match_bowkf.npz")
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            cnt, m = R.ref_search_by_bow_keyframes(bow_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3])), float(c[4]), bool(c[5]))
            assert cnt == int(g["n_%d" % i]) and np.array_equal(m, g["m_%d" % i]), i
        g = gold("ref_match_triangulation.npz")
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            w = triangulation_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3]), stereo_fraction=float(c[4]), forward=bool(c[5]))
            cnt, m, _ = R.ref_search_for_triangulation(w, bool(c[6]), bool(c[7]))
            assert cnt == int(g["n_%d" % i]) and np.array_equal(m, g["m_%d" % i]), i


def test_fuse_sim3_and_projection_sim3_shim_reproduce_reference():
    """Fuse: the harness calls ORBmatcher::Fuse once per candidate on an empty key frame and reads the chosen keypoint
    back from pMP->mObservations, so the shim's host-side add-observation surgery is executed as well."""
    with R.shim_bodies():
        g = gold("ref_match_fuse.npz")
        c = g["cfg_2"]                                                     # 800 candidates: one device round trip each
        assert np.array_equal(R.ref_fuse_search(fuse_frame(int(c[0]), int(c[1]), int(c[2])), BOUNDS, float(c[3])), g["best_2"])
        g = gold("ref_match_sim3.npz")
        w = fuse_frame(885, 600, 900)
        assert np.array_equal(R.ref_fuse_search_sim3(w, BOUNDS, 4.0), O.fuse_search(w, BOUNDS, 4.0, 1)[0])
        for i in range(int(g["scount"])):
            c = g["scfg_%d" % i]
            found, m = R.ref_search_by_sim3(sim3_pair(int(c[0]), int(c[1]), int(c[2])), BOUNDS, float(c[3]))
            assert found == int(g["sn_%d" % i]) and np.array_equal(m, g["sm_%d" % i]), i
        g = gold("ref_match_projsim3.npz")
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            pre = np.where(np.random.default_rng(int(c[0])).random(int(c[2])) < 0.1, -2, -1).astype(np.int32)
            cnt, m = R.ref_search_by_projection_sim3(fuse_frame(int(c[0]), int(c[1]), int(c[2])), BOUNDS, int(c[3]), pre)
            assert cnt == int(g["n_%d" % i]) and np.array_equal(m, g["m_%d" % i]), i


def test_distinctive_descriptors_shim():
    obs = observed_descriptors(4, [1, 2, 3, 5, 8, 17, 32, 33, 64, 100])
    with R.shim_bodies():
        for d in obs:
            idx, _ = O.distinctive_descriptor(d)
            got = R.ref_compute_distinctive_descriptors(d)
            assert got is not None and np.array_equal(got, d[idx])


def test_frame_compute_bow_shim_reproduces_reference(tmp_path):
    g = gold("ref_bow_transform.npz")
    L = R.shimlib()
    f = L.shimm_frame_compute_bow
    f.argtypes = [C.c_char_p, C.c_int, _u8p, C.POINTER(C.c_int32), O._u32p, O._f64p, C.POINTER(C.c_int32), O._u32p, _i32p, O._u32p]
    f.restype = C.c_int
    done = 0
    for i in range(int(g["count"])):
        vi, k, Lv, n, lu = (int(v) for v in g["cfg_%d" % i])
        if lu != 4:                                                        # Frame::ComputeBoW always passes levelsup = 4
            continue
        voc = synthetic_vocabulary(vi, k, Lv)
        path = os.path.join(str(tmp_path), "voc%d.txt" % vi)
        write_vocabulary_text(path, voc)
        d = O._b(vocabulary_features(vi, voc, n)).reshape(-1, 32)
        o = O._bow_out(len(d))
        rc = f(path.encode(), len(d), O._ptr(d, _u8p), C.byref(o["bn"]), O._ptr(o["bw"], O._u32p), O._ptr(o["bv"], O._f64p),
               C.byref(o["fn"]), O._ptr(o["fnode"], O._u32p), O._ptr(o["fstart"], _i32p), O._ptr(o["ffeat"], O._u32p))
        assert rc == 0
        r = O._bow_result(o)
        for key in ("word", "node", "start", "feat"):
            assert np.array_equal(r[key], g["%s_%d" % (key, i)]), (i, key)
        assert r["value"].tobytes() == g["value_%d" % i].tobytes()
        done += 1
    assert done >= 2


def test_compute_stereo_matches_shim_host_pyramids_reproduces_reference():
    """Frame::ComputeStereoMatches (shim body) on a Frame whose extractors only carry host pyramids (mvImagePyramid)."""
    g = gold("ref_stereo_matches.npz")
    idx, w, h, nf, mb, mbf = g["cfg_0"]
    left = synthetic_frame(int(idx), int(w), int(h)); right = stereo_right_frame(left, int(idx))
    side = []
    for img in (left, right):
        ex = O.OracleExtractor(int(nf), 1.2, 8, 20, 7)
        k, d = ex(img)
        side.append((k, d, [ex.level_pixels(l) for l in range(8)], ex.scale_factors, ex.inv_scale_factors))
    (kl, dl, lp, sc, isc), (kr, dr, rp, _, _) = side
    with R.shim_bodies():
        ur, dep, cnt = R.ref_compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, float(mb), float(mbf))
    assert cnt == int(g["n_0"]) and ur.tobytes() == g["ur_0"].tobytes() and dep.tobytes() == g["depth_0"].tobytes()


def _frames_init(L, a, b, nfeatures, K, dist, ratio=0.9, window=100):
    f = L.shimm_rgbd_frames_init_match
    f.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_int,
                  C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, _u8p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, _u8p,
                  _f32p, _i32p, _f32p]
    f.restype = C.c_int
    cap = nfeatures + 64
    h, w = a.shape
    out = {}
    for s in "AB":
        out["n" + s] = C.c_int32(0)
        out["k" + s] = np.zeros(cap, O.KP_DTYPE); out["u" + s] = np.zeros(cap, O.KP_DTYPE); out["d" + s] = np.zeros((cap, 32), np.uint8)
    bounds = np.zeros(4, np.float32); m12 = np.full(cap, -1, np.int32); prev = np.zeros((cap, 2), np.float32)
    K = O._f(K); dist = O._f(dist)
    n = f(O._ptr(a, _u8p), O._ptr(b, _u8p), w, h, nfeatures, O._ptr(K, _f32p), O._ptr(dist, _f32p), ratio, window, cap,
          C.byref(out["nA"]), out["kA"].ctypes.data, out["uA"].ctypes.data, O._ptr(out["dA"], _u8p),
          C.byref(out["nB"]), out["kB"].ctypes.data, out["uB"].ctypes.data, O._ptr(out["dB"], _u8p),
          O._ptr(bounds, _f32p), O._ptr(m12, _i32p), O._ptr(prev, _f32p))
    return n, out, bounds, m12, prev


@pytest.mark.parametrize("distorted", [False, True])
def test_reference_frame_constructor_and_matcher_on_shim_classes(distorted):
    """The reference's own RGB-D Frame constructor (S/Frame.cc:135-188) run on the drop-in ORBextractor, then
    SearchForInitialization as Tracking::MonocularInitialization calls it: extraction, UndistortKeyPoints,
    ComputeImageBounds, AssignFeaturesToGrid (reference body) and the matcher, against the oracle chain."""
    from weiner_slamit_v2_b200.pipeline import REFERENCE_DIST, REFERENCE_K
    dist = np.asarray(REFERENCE_DIST if distorted else [0, 0, 0, 0, 0], np.float32)
    a = synthetic_frame(60); b = np.ascontiguousarray(np.roll(a, (2, 4), (0, 1)))
    n, out, bounds, m12, prev = _frames_init(R.shimlib(), a, b, 1000, REFERENCE_K, dist)
    orc = O.OracleExtractor()
    k1, d1 = orc(a); k2, d2 = orc(b)
    assert out["nA"].value == len(k1) and out["nB"].value == len(k2)
    assert out["kA"][:len(k1)].tobytes() == k1.tobytes() and np.array_equal(out["dA"][:len(k1)], d1)
    assert out["kB"][:len(k2)].tobytes() == k2.tobytes() and np.array_equal(out["dB"][:len(k2)], d2)
    ob = O.image_bounds(640, 480, REFERENCE_K, dist)
    assert np.array_equal(bounds, ob)
    if distorted:
        for k in (k1, k2):
            u = O.undistort_points(np.stack([k["x"], k["y"]], 1), REFERENCE_K, dist)
            k["x"], k["y"] = u[:, 0], u[:, 1]
    assert out["uA"][:len(k1)].tobytes() == k1.tobytes() and out["uB"][:len(k2)].tobytes() == k2.tobytes()
    p0 = np.stack([k1["x"], k1["y"]], 1)
    on, om12, opm = O.search_for_initialization(k1, d1, k2, d2, p0, ob, 0.9, True, 100)
    assert n == on and on > 20
    assert np.array_equal(m12[:len(k1)], om12) and np.array_equal(prev[:len(k1)], opm)


def test_reference_stereo_frame_constructor_on_shim_classes():
    """The reference's stereo Frame constructor (S/Frame.cc:70-133): two drop-in extractors driven from the two
    std::threads of :93-96, then Frame::ComputeStereoMatches on the pyramids they left on the device."""
    g = gold("ref_stereo_matches.npz")
    L = R.shimlib()
    f = L.shimm_stereo_frame
    f.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _f32p, C.c_float, C.c_int, C.POINTER(C.c_int32), C.c_void_p, _u8p,
                  C.POINTER(C.c_int32), C.c_void_p, _u8p, _f32p, _f32p]
    f.restype = C.c_int
    for i in range(int(g["count"])):
        idx, w, h, nf, mb, mbf = g["cfg_%d" % i]
        left = synthetic_frame(int(idx), int(w), int(h)); right = stereo_right_frame(left, int(idx))
        cap = int(nf) + 64
        nL, nR = C.c_int32(0), C.c_int32(0)
        kL = np.zeros(cap, O.KP_DTYPE); kR = np.zeros(cap, O.KP_DTYPE)
        dL = np.zeros((cap, 32), np.uint8); dR = np.zeros((cap, 32), np.uint8)
        ur = np.zeros(cap, np.float32); dep = np.zeros(cap, np.float32)
        K = np.array([400.0, 400.0, w / 2, h / 2], np.float32)
        cnt = f(O._ptr(left, _u8p), O._ptr(right, _u8p), int(w), int(h), int(nf), O._ptr(K, _f32p), float(mbf), cap,
                C.byref(nL), kL.ctypes.data, O._ptr(dL, _u8p), C.byref(nR), kR.ctypes.data, O._ptr(dR, _u8p), O._ptr(ur, _f32p), O._ptr(dep, _f32p))
        n = nL.value
        assert np.array_equal(_sha(kL[:n]), g["kl_sha_%d" % i]) and np.array_equal(_sha(kR[:nR.value]), g["kr_sha_%d" % i])
        # the constructor runs ComputeStereoMatches before it assigns mb (S/Frame.cc:104 against :130): mb = 0 in the
        # harness's zeroed Frame, i.e. maxD = +inf -- the oracle with the same value
        pyr = []
        for img in (left, right):
            ex = O.OracleExtractor(int(nf), 1.2, 8, 20, 7)
            ex(img)
            pyr.append(([ex.level_pixels(l) for l in range(8)], ex.scale_factors, ex.inv_scale_factors))
        (lp, sc, isc), (rp, _, _) = pyr
        with np.errstate(divide="ignore"):
            ur_o, dep_o, kept, skipped = O.compute_stereo_matches(kL[:n], dL[:n], kR[:nR.value], dR[:nR.value], sc, isc, lp, rp, 0.0, float(mbf))
        assert cnt == kept and cnt > 50 and ur[:n].tobytes() == ur_o.tobytes() and dep[:n].tobytes() == dep_o.tobytes(), i


def test_tracking_extractor_sequence_on_one_thread():
    """Tracking.cc:156-162 creates mpIniORBextractor (2 * nFeatures) and mpORBextractorLeft (nFeatures); the first is used
    until initialisation, then the second, and Reset() returns to the first: handles of different sizes in turn."""
    L = R.shimlib()
    f = L.shimm_extractor_sequence
    f.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p]
    img = synthetic_frame(11)
    counts = np.zeros(3, np.int32)
    assert f(O._ptr(img, _u8p), 640, 480, 1000, O._ptr(counts, _i32p)) == 0
    big = len(O.OracleExtractor(2000, 1.2, 8, 20, 7)(img)[0]); small = len(O.OracleExtractor(1000, 1.2, 8, 20, 7)(img)[0])
    assert list(counts) == [big, small, big]


def test_matcher_shim_from_three_host_threads():
    """ORBmatcher is used concurrently from the Tracking, LocalMapping and LoopClosing threads (S/System.cc:156,160):
    three host threads call different entry points of the drop-in class at once (ctypes releases the GIL), each through
    its own thread-local device handle.  All three use ONE image rectangle, as one camera does: Frame::mnMinX ... are
    static members of the reference's Frame, which every harness call sets."""
    common = (0, 0, 1280, 720)
    gp = gold("ref_match_proj.npz")
    R.shimlib()
    errors = []
    p = init_pair(400, n=1000)
    want_init = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], common, 0.9, True, 100)
    c = gp["cfg_1"]
    kp, kd, mp = projection_frame(int(c[0]), int(c[1]), int(c[2]))
    w = motion_frame(700)
    want_last = O.search_by_projection_last_frame(w, SCALE_FACTORS_8, common, 15.0, 0, True)
    assert want_init[0] > 20 and want_last[0] > 100

    def init_worker():
        for _ in range(6):
            cnt, m12, prev = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], p[4], common, 0.9, True, 100)
            if cnt != want_init[0] or not np.array_equal(m12, want_init[1]) or not np.array_equal(prev, want_init[2]):
                errors.append("init")

    def proj_worker():
        for _ in range(6):
            cnt, kpmp = R.ref_search_by_projection(mp, kp, kd, SCALE_FACTORS_8, common, 0.8, float(c[3]))
            if cnt != int(gp["n_1"]) or not np.array_equal(kpmp, gp["kpmp_1"]):
                errors.append("proj")

    def last_worker():
        for _ in range(6):
            cnt, kpmp = R.ref_search_by_projection_last_frame(w, SCALE_FACTORS_8, common, 15.0, True)
            if cnt != want_last[0] or not np.array_equal(kpmp, want_last[1]):
                errors.append("last")

    with R.shim_bodies():
        threads = [threading.Thread(target=t) for t in (init_worker, proj_worker, last_worker)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    assert errors == []
