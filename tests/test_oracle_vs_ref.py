"""Pins the C oracle against the reference's OWN code: oracle/_ref = ORB_SLAM2/src/ORBextractor.cc
compiled unmodified for x86 (built here by `make -C oracle ref`; prebuilt file on the GPU box), and
against the golden vectors that build produced (tests/golden/, tools/gen_golden.py)."""
import glob
import hashlib
import os

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R
from weiner_slamit_v2_b200 import frames as F

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_extract_*.npz")))
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference)")


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_golden_vectors(path):
    g = np.load(path)
    p = g["params"]
    params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))
    meta = [str(m) for m in g["meta"]]
    gen = getattr(F, meta[0])
    w, h = int(meta[1]), int(meta[2])
    orc = O.OracleExtractor(*params)
    for i, idx in enumerate(meta[3:]):
        img = gen(int(idx), w, h)
        assert np.array_equal(_sha(img), g["frame_sha_%d" % i]), "frame generator changed"
        k, d = orc(img)
        assert k.tobytes() == g["kps_%d" % i].tobytes()
        assert np.array_equal(d, g["desc_%d" % i])
        for l in range(params[2]):
            assert np.array_equal(_sha(orc.level_pixels(l)), g["pyr_sha_%d" % i][l])


def test_golden_vectors_exist():
    assert len(GOLDEN) >= 4


@needs_ref
def test_constructor_tables_match_reference():
    for params in ((1000, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (500, 1.5, 5, 25, 9), (1500, 1.1, 12, 20, 7)):
        t = R.RefExtractor(*params).tables()
        o = O.OracleExtractor(*params)
        assert np.array_equal(t["scale"], o.scale_factors) and np.array_equal(t["inv_scale"], o.inv_scale_factors)
        assert np.array_equal(t["sigma2"], o.level_sigma2) and np.array_equal(t["inv_sigma2"], o.inv_level_sigma2)
        assert np.array_equal(t["per_level"], o.features_per_level) and np.array_equal(t["umax"], o.umax)
        assert np.array_equal(t["pattern"], np.ctypeslib.as_array(O.lib().orc_pattern(), (1024,)))


@needs_ref
def test_full_extractor_matches_reference_fresh_frames():
    o, r = O.OracleExtractor(), R.RefExtractor()
    for i in (11, 12, 13):
        img = F.synthetic_frame(i)
        ko, do = o(img)
        kr, dr = r(img)
        assert ko.tobytes() == kr.tobytes() and np.array_equal(do, dr)
        for l in range(8):
            assert np.array_equal(o.level_pixels(l), r.level_pixels(l))
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (480, 640)).astype(np.uint8)      # far more candidates than the quota
    ko, do = o(img); kr, dr = r(img)
    assert ko.tobytes() == kr.tobytes() and np.array_equal(do, dr)
    assert len(o(np.full((480, 640), 9, np.uint8))[0]) == len(r(np.full((480, 640), 9, np.uint8))[0]) == 0


@needs_ref
def test_retry_when_nms_empties_a_cell_matches_reference():
    """FAST(20) has corners in the cell but NMS keeps none (a plateau of equal scores) -> the reference retries at 7
    (ORBextractor.cc:827-833).  The frames must really contain such cells: candidates below 20 next to corners at 20."""
    o, r = O.OracleExtractor(), R.RefExtractor()
    for i in range(6):
        img = F.plateau_retry_frame(i) if i < 3 else F.plateau_retry_frame(i, 320, 240)
        ko, do = o(img)
        kr, dr = r(img)
        assert len(ko) > 0 and ko.tobytes() == kr.tobytes() and np.array_equal(do, dr)
        resp = o.level_candidates(0)["response"]
        assert (resp < 20).any() and (resp >= 20).any()


@needs_ref
def test_pyramid_border_matches_reference():
    r = R.RefExtractor(); o = O.OracleExtractor()
    img = F.synthetic_frame(5)
    r(img); o(img)
    for l in range(8):
        assert np.array_equal(r.level_pixels(l, with_border=True), O.copy_make_border(o.level_pixels(l), 19))


@needs_ref
def test_distribute_octree_matches_reference_under_monotonic_allocator():
    """The quadtree's (count, node address) tie-break (S/ORBextractor.cc:694-698) is fixed by the bump
    allocator in oracle/ref_harness.cc; the oracle uses creation order, which must agree."""
    r = R.RefExtractor()
    rng = np.random.default_rng(21)
    for trial in range(40):
        w, h = int(rng.integers(120, 1300)), int(rng.integers(100, 700))
        if round(w / h) < 1:
            continue
        n = int(rng.integers(0, 3000))
        N = int(rng.integers(1, 500))
        pts = set()
        while len(pts) < n:
            pts.add((int(rng.integers(3, w - 3)), int(rng.integers(3, h - 3))))
        c = np.zeros(len(pts), O.KP_DTYPE)
        arr = np.array(sorted(pts, key=lambda p: (p[1], p[0])), np.float32).reshape(-1, 2)
        if len(arr):
            c["x"], c["y"] = arr[:, 0], arr[:, 1]
        c["response"] = rng.integers(7, 60, len(c)).astype(np.float32)   # many response ties
        c["size"] = 7; c["angle"] = -1; c["class_id"] = -1
        a = O.distribute_octree(c, 16, 16 + w, 16, 16 + h, N)
        b = r.distribute_octree(c, 16, 16 + w, 16, 16 + h, N)
        assert a.tobytes() == b.tobytes(), (trial, w, h, n, N, len(a), len(b))


# ---- matcher: oracle/_ref/libref_matcher.so = the reference's own ORBmatcher.cc + Frame.cc + MapPoint.cc ------
needs_refm = pytest.mark.skipif(not R.matcher_available(), reason="oracle/_ref/libref_matcher.so not built")
BOWGOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_bow_transform.npz")
MGOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_match_*.npz")))


@needs_refm
def test_descriptor_distance_matches_reference():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (500, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (500, 32)).astype(np.uint8)
    b[:5] = a[:5]; b[5:10] = ~a[5:10]
    for i in range(500):
        assert O.descriptor_distance(a[i], b[i]) == R.ref_descriptor_distance(a[i], b[i])


@needs_refm
@pytest.mark.parametrize("brute,window,n,ori", [(False, 100, 1000, True), (True, 1000, 1000, True), (False, 10, 700, True),
                                                 (True, 40, 333, False), (False, 100, 1, True), (False, 60, 0, True)])
def test_search_for_initialization_matches_reference(brute, window, n, ori):
    from weiner_slamit_v2_b200.workloads import init_pair
    for idx in range(3):
        p = init_pair(200 + idx, n=max(n, 1), brute_force=brute)
        if n == 0:
            p = tuple(x[:0] for x in p)
        for bounds in ((0, 0, 640, 480), (-13.7, -9.2, 661.3, 492.8)):
            a = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], bounds, 0.9, ori, window)
            b = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], p[4], bounds, 0.9, ori, window)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
            # second round from the updated vbPrevMatched, as MonocularInitialization does
            a2 = O.search_for_initialization(p[0], p[1], p[2], p[3], a[2], bounds, 0.9, ori, window)
            b2 = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], b[2], bounds, 0.9, ori, window)
            assert a2[0] == b2[0] and np.array_equal(a2[1], b2[1]) and np.array_equal(a2[2], b2[2])


@needs_refm
@pytest.mark.parametrize("th", [1.0, 3.0, 5.0])
def test_search_by_projection_matches_reference(th):
    from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, projection_frame
    for idx, (w, h) in enumerate(((1280, 720), (640, 480))):
        kp, kd, mp = projection_frame(300 + idx, 2000, 4000, w, h)
        rng = np.random.default_rng(idx)
        pre = np.full(2000, -1, np.int32); obs = np.zeros(2000, np.int32)
        ii = rng.choice(2000, 150, replace=False)
        pre[ii] = -2; obs[ii] = rng.integers(0, 3, 150)
        pre[rng.choice(2000, 50, replace=False)] = rng.integers(0, 4000, 50)     # some already hold listed map points
        ur = np.where(rng.random(2000) < 0.3, kp["x"] - rng.uniform(0, 30, 2000), -1).astype(np.float32)
        mp["kuright"] = ur
        mp["xr"] = (mp["x"] - rng.uniform(0, 30, 4000)).astype(np.float32)
        a = O.search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, w, h), 0.8, th, pre, obs)
        b = R.ref_search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, w, h), 0.8, th, pre, obs)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("path", MGOLDEN, ids=[os.path.basename(p) for p in MGOLDEN])
def test_matcher_oracle_reproduces_reference_golden_vectors(path):
    """Vectors produced by the reference's own matcher (tools/gen_golden.py); checked wherever the tests run."""
    from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, init_pair, projection_frame
    g = np.load(path)
    if "projsim3" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import fuse_frame
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            pre = np.where(np.random.default_rng(int(c[0])).random(int(c[2])) < 0.1, -2, -1).astype(np.int32)
            a = O.search_by_projection_sim3(fuse_frame(int(c[0]), int(c[1]), int(c[2])), (-13.7, -9.2, 661.3, 492.8), int(c[3]), pre)
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["m_%d" % i])
    elif "sim3" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import fuse_frame, sim3_pair
        bounds = (-13.7, -9.2, 661.3, 492.8)
        for i in range(int(g["fcount"])):
            c = g["fcfg_%d" % i]
            a = O.fuse_search(fuse_frame(int(c[0]), int(c[1]), int(c[2])), bounds, float(c[3]), 1)
            assert np.array_equal(a[0], g["fbest_%d" % i])
        for i in range(int(g["scount"])):
            c = g["scfg_%d" % i]
            a = O.search_by_sim3(sim3_pair(int(c[0]), int(c[1]), int(c[2])), bounds, float(c[3]))
            assert a[0] == int(g["sn_%d" % i]) and np.array_equal(a[1], g["sm_%d" % i])
    elif "fuse" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import fuse_frame
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            a = O.fuse_search(fuse_frame(int(c[0]), int(c[1]), int(c[2])), (-13.7, -9.2, 661.3, 492.8), float(c[3]))
            assert np.array_equal(a[0], g["best_%d" % i])
    elif "triangulation" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import triangulation_pair
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            w = triangulation_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3]), stereo_fraction=float(c[4]), forward=bool(c[5]))
            a = O.search_for_triangulation(w, bool(c[6]), bool(c[7]))
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["m_%d" % i])
    elif "bow" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import bow_pair
        fn = O.search_by_bow_keyframes if "bowkf" in os.path.basename(path) else O.search_by_bow
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            a = fn(bow_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3])), float(c[4]), bool(c[5]))
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["m_%d" % i])
    elif "keyframe" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import relocalisation_frame
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            a = O.search_by_projection_keyframe(relocalisation_frame(int(c[0])), SCALE_FACTORS_8, (-13.7, -9.2, 661.3, 492.8),
                                                float(c[1]), int(c[2]), bool(c[3]))
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["kpmp_%d" % i])
    elif "lastframe" in os.path.basename(path):
        from weiner_slamit_v2_b200.workloads import motion_frame
        for i in range(int(g["count"])):
            idx, th, ori = int(g["cfg_%d" % i][0]), float(g["cfg_%d" % i][1]), bool(g["cfg_%d" % i][2])
            a = O.search_by_projection_last_frame(motion_frame(idx), SCALE_FACTORS_8, (-13.7, -9.2, 661.3, 492.8), th, 0, ori)
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["kpmp_%d" % i])
    elif "init" in os.path.basename(path):
        for i in range(int(g["count"])):
            idx, n, brute, window = (int(v) for v in g["cfg_%d" % i])
            p = init_pair(idx, n=n, brute_force=bool(brute))
            a = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, window)
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["m12_%d" % i]) and np.array_equal(a[2], g["prev_%d" % i])
    else:
        for i in range(int(g["count"])):
            idx, nk, nmp = (int(v) for v in g["cfg_%d" % i][:3])
            th = float(g["cfg_%d" % i][3])
            kp, kd, mp = projection_frame(idx, nk, nmp)
            a = O.search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, th)
            assert a[0] == int(g["n_%d" % i]) and np.array_equal(a[1], g["kpmp_%d" % i])


def test_matcher_golden_vectors_exist():
    assert len(MGOLDEN) >= 3


@needs_refm
@pytest.mark.parametrize("th,ori", [(7.0, True), (15.0, True), (15.0, False)])
def test_search_by_projection_last_frame_matches_reference(th, ori):
    """Scope row N2 against the reference's own code (its matrix expressions evaluated by oracle/cv341_stubs.cc,
    whose float arithmetic is pinned to cv2.gemm in test_oracle_primitives.py)."""
    from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, motion_frame
    bounds = (-13.7, -9.2, 661.3, 492.8)
    for idx in range(4):
        w = motion_frame(600 + idx)
        rng = np.random.default_rng(idx)
        n = len(w["cur"])
        pre = np.full(n, -1, np.int32); obs = np.zeros(n, np.int32)
        ii = rng.choice(n, 100, replace=False)
        pre[ii[:50]] = -2; obs[ii[:50]] = rng.integers(0, 3, 50)
        pre[ii[50:]] = rng.integers(0, len(w["has_mp"]), 50)
        a = O.search_by_projection_last_frame(w, SCALE_FACTORS_8, bounds, th, 0, ori, 40.0, pre, obs)
        b = R.ref_search_by_projection_last_frame(w, SCALE_FACTORS_8, bounds, th, ori, 40.0, pre, obs)
        assert a[0] == b[0] and np.array_equal(a[1], b[1])


@needs_refm
@pytest.mark.parametrize("th,orb_dist,ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_keyframe_matches_reference(th, orb_dist, ori):
    """The relocalisation overload (S/ORBmatcher.cc:1476-1603) against the reference's own code, including
    MapPoint::PredictScale (logf) and cv::norm; Ow as the harness evaluates -Rcw.t()*tcw equals the workload's."""
    from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, relocalisation_frame
    for idx, bounds in enumerate(((0.0, 0.0, 640.0, 480.0), (-13.7, -9.2, 661.3, 492.8), (0.0, 0.0, 640.0, 480.0))):
        w = relocalisation_frame(900 + idx, n_kf=1500 if idx < 2 else 40, n_cur=2000 if idx < 2 else 60)
        a = O.search_by_projection_keyframe(w, SCALE_FACTORS_8, bounds, th, orb_dist, ori)
        b = R.ref_search_by_projection_keyframe(w, SCALE_FACTORS_8, bounds, th, orb_dist, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(b[2], w["Ow"])
        assert idx == 2 or a[0] > 100


@needs_refm
@pytest.mark.parametrize("ratio,ori", [(0.7, True), (0.9, True), (0.7, False)])
def test_search_by_bow_matches_reference(ratio, ori):
    """SearchByBoW(KeyFrame*, Frame&, ...) (S/ORBmatcher.cc:161-292) against the reference's own code, the feature
    vectors rebuilt with the reference's DBoW2::FeatureVector::addFeature: many nodes, few nodes, a single node,
    more nodes than features, empty sides."""
    from weiner_slamit_v2_b200.workloads import bow_pair
    for idx, (nk, nf, nn) in enumerate([(2000, 2000, 100), (2000, 1500, 100), (500, 800, 30), (0, 100, 10), (100, 0, 10),
                                        (50, 50, 1), (1000, 1000, 1000), (300, 300, 2)]):
        w = bow_pair(960 + idx, nk, nf, nn)
        a = O.search_by_bow(w, ratio, ori)
        b = R.ref_search_by_bow(w, ratio, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), idx
        a = O.search_by_bow_keyframes(w, ratio, ori)                # SearchByBoW(pKF1, pKF2, ...) (:526-659)
        b = R.ref_search_by_bow_keyframes(w, ratio, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), idx


@needs_refm
@pytest.mark.parametrize("only_stereo,ori", [(False, True), (True, True), (False, False)])
def test_search_for_triangulation_matches_reference(only_stereo, ori):
    """SearchForTriangulation (S/ORBmatcher.cc:661-827) incl. CheckDistEpipolarLine against the reference's own code:
    sideways and forward motion (epipole inside the image), mono and partly stereo key frames, empty sides.  The
    epipole the harness computes with the reference's expressions equals the workload's."""
    from weiner_slamit_v2_b200.workloads import triangulation_pair
    cfg = [(2000, 2000, 100, 0.0, False), (2000, 1500, 100, 0.3, False), (2000, 2000, 100, 0.0, True), (1200, 1200, 50, 0.3, True),
           (0, 100, 10, 0.0, False), (100, 0, 10, 0.0, False), (50, 50, 1, 0.5, False), (1000, 1000, 1000, 0.0, True)]
    tot = 0
    for idx, (n1, n2, nn, sfr, fwd) in enumerate(cfg):
        w = triangulation_pair(980 + idx, n1, n2, nn, stereo_fraction=sfr, forward=fwd)
        a = O.search_for_triangulation(w, only_stereo, ori)
        b = R.ref_search_for_triangulation(w, only_stereo, ori)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(b[2], w["epipole"]), idx
        tot += a[0]
    assert tot > (20 if only_stereo else 400)


@needs_refm
def test_distinctive_descriptor_matches_reference():
    """MapPoint::ComputeDistinctiveDescriptors (S/MapPoint.cc:248-313) run natively on fake observations (some key
    frames bad): the oracle picks the same descriptor for 1..200 observations."""
    from weiner_slamit_v2_b200.workloads import observed_descriptors
    rng = np.random.default_rng(3)
    sizes = [1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 50, 100, 200] * 3
    for k, d in enumerate(observed_descriptors(0, sizes)):
        bad = (rng.random(len(d)) < 0.15).astype(np.uint8)
        good = d[bad == 0]
        idx, med = O.distinctive_descriptor(good)
        r = R.ref_compute_distinctive_descriptors(d, bad)
        if len(good) == 0:
            assert r is None and idx == -1
        else:
            assert np.array_equal(r, good[idx]), k


def test_distinctive_descriptor_small_cases():
    d = np.zeros((3, 32), np.uint8); d[1, 0] = 0xff; d[2, :2] = 0xff          # distances 0-1: 8, 0-2: 16, 1-2: 8
    assert O.distinctive_descriptor(d) == (0, 8)                               # every row's median is 8: the first row wins
    d = np.zeros((4, 32), np.uint8); d[0, :4] = 0xff; d[1, 0] = 0x0f; d[3, 0] = 0x03
    assert O.distinctive_descriptor(d) == (1, 2)                               # rows sorted: (0,28,30,32) (0,2,4,28) (0,2,4,32) (0,2,2,30) -> element 1
    assert O.distinctive_descriptor(np.zeros((0, 32), np.uint8))[0] == -1
    assert O.distinctive_descriptor(np.full((1, 32), 7, np.uint8)) == (0, 0)


@needs_refm
@pytest.mark.parametrize("th", [2.5, 3.0, 10.0])
def test_fuse_search_matches_reference(th):
    """The search inside Fuse(pKF, vpMapPoints, th) (S/ORBmatcher.cc:829-948) against the reference's own Fuse called
    once per candidate on an empty key frame (oracle/ref_matcher_harness.cc): integer and fractional image bounds
    (the key frame truncates them), mono and stereo keypoints, NULL / bad / already-observed candidates."""
    from weiner_slamit_v2_b200.workloads import fuse_frame
    tot = 0
    for idx, (nmp, nkp, bounds, sfr) in enumerate([(3000, 2000, (0.0, 0.0, 640.0, 480.0), 0.2), (3000, 2000, (-13.7, -9.2, 661.3, 492.8), 0.0),
                                                   (500, 300, (-13.7, -9.2, 661.3, 492.8), 0.5), (0, 100, (0.0, 0.0, 640.0, 480.0), 0.2),
                                                   (100, 0, (0.0, 0.0, 640.0, 480.0), 0.2)]):
        w = fuse_frame(990 + idx, nmp, nkp, stereo_fraction=sfr)
        a = O.fuse_search(w, bounds, th)
        b = R.ref_fuse_search(w, bounds, th)
        assert np.array_equal(a[0], b), idx
        tot += int((b >= 0).sum())
    assert tot > 400


@needs_refm
def test_fuse_sim3_search_and_search_by_sim3_match_reference():
    """Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (S/ORBmatcher.cc:979-1104; one candidate per call, see the harness)
    and SearchBySim3 (:1106-1330; both legs and the agreement test in one call of the reference) against the oracle's
    mode 1 / mode 2 searches."""
    from weiner_slamit_v2_b200.workloads import fuse_frame, sim3_pair
    tot = 0
    for idx, (nmp, nkp, bounds) in enumerate([(3000, 2000, (0.0, 0.0, 640.0, 480.0)), (3000, 2000, (-13.7, -9.2, 661.3, 492.8)),
                                              (500, 300, (-13.7, -9.2, 661.3, 492.8)), (0, 100, (0.0, 0.0, 640.0, 480.0))]):
        for th in (3.0, 4.0):
            w = fuse_frame(1000 + idx, nmp, nkp)
            a = O.fuse_search(w, bounds, th, 1)
            b = R.ref_fuse_search_sim3(w, bounds, th)
            assert np.array_equal(a[0], b), idx
            tot += int((b >= 0).sum())
    assert tot > 800
    tot = 0
    for idx, (n1, n2, bounds) in enumerate([(1500, 1500, (0.0, 0.0, 640.0, 480.0)), (1500, 1200, (-13.7, -9.2, 661.3, 492.8)),
                                            (300, 400, (0.0, 0.0, 640.0, 480.0)), (0, 100, (0.0, 0.0, 640.0, 480.0)), (100, 0, (0.0, 0.0, 640.0, 480.0))]):
        for th in (7.5, 10.0):
            w = sim3_pair(1010 + idx, n1, n2)
            a = O.search_by_sim3(w, bounds, th)
            b = R.ref_search_by_sim3(w, bounds, th)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]), idx
            tot += b[0]
    assert tot > 300


@needs_refm
@pytest.mark.parametrize("th", [4, 10])
def test_search_by_projection_sim3_matches_reference(th):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (S/ORBmatcher.cc:294-407) run whole by the reference,
    with pre-occupied keypoints and more candidates than keypoints (the greedy occupancy matters)."""
    from weiner_slamit_v2_b200.workloads import fuse_frame
    rng = np.random.default_rng(4)
    tot = 0
    for idx, (nmp, nkp, bounds) in enumerate([(3000, 2000, (0.0, 0.0, 640.0, 480.0)), (3000, 2000, (-13.7, -9.2, 661.3, 492.8)),
                                              (500, 300, (-13.7, -9.2, 661.3, 492.8)), (0, 100, (0.0, 0.0, 640.0, 480.0)),
                                              (100, 0, (0.0, 0.0, 640.0, 480.0)), (6000, 1000, (0.0, 0.0, 640.0, 480.0))]):
        w = fuse_frame(1020 + idx, nmp, nkp)
        pre = np.where(rng.random(nkp) < 0.1, -2, -1).astype(np.int32)
        a = O.search_by_projection_sim3(w, bounds, th, pre)
        b = R.ref_search_by_projection_sim3(w, bounds, th, pre)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]), idx
        tot += b[0]
    assert tot > 800


def _same_bow(a, b):
    return all(np.array_equal(a[x], b[x]) for x in ("word", "node", "start", "feat")) and a["value"].tobytes() == np.asarray(b["value"]).tobytes()


@needs_refm
def test_bow_transform_matches_reference(tmp_path):
    """Frame::ComputeBoW's DBoW2 transform against the reference's own TemplatedVocabulary (vocabulary written in
    ORBvoc.txt format and read back by the reference's loadFromTextFile): words, bit-identical tf-idf values, feature
    vector; levelsup above the tree height (root node), a single feature, no features, stopped words."""
    from weiner_slamit_v2_b200.workloads import synthetic_vocabulary, vocabulary_features, write_vocabulary_text
    for vi, (k, L) in enumerate([(6, 4), (10, 3), (3, 6), (4, 2)]):
        voc = synthetic_vocabulary(20 + vi, k, L)
        path = str(tmp_path / ("voc%d.txt" % vi))
        write_vocabulary_text(path, voc)
        for n, lu in [(2000, 4), (500, 2), (1, 4), (0, 4), (1000, 1), (1000, 10)]:
            d = vocabulary_features(vi * 10 + n % 7, voc, n)
            assert _same_bow(O.bow_transform(voc, d, lu), R.ref_bow_transform(path, d, lu)), (vi, n, lu)


def test_bow_transform_oracle_reproduces_reference_golden_vectors():
    from weiner_slamit_v2_b200.workloads import synthetic_vocabulary, vocabulary_features
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_bow_transform.npz"))
    for i in range(int(g["count"])):
        vi, k, L, n, lu = (int(v) for v in g["cfg_%d" % i])
        voc = synthetic_vocabulary(vi, k, L)
        a = O.bow_transform(voc, vocabulary_features(vi, voc, n), lu)
        assert _same_bow(a, {key: g["%s_%d" % (key, i)] for key in ("word", "value", "node", "start", "feat")})


# ---- Frame::ComputeStereoMatches -----------------------------------------------------------------------------------
def _stereo_case(index, w=640, h=480, nfeat=1000):
    left = F.synthetic_frame(index, w, h)
    right = F.stereo_right_frame(left, index)
    out = []
    for img in (left, right):
        ex = O.OracleExtractor(nfeat, 1.2, 8, 20, 7)
        k, d = ex(img)
        out.append((k, d, [ex.level_pixels(l) for l in range(8)], ex.scale_factors, ex.inv_scale_factors))
    return out


@needs_refm
@pytest.mark.parametrize("index,mb,mbf", [(0, 0.1, 40.0), (1, 0.5, 20.0), (2, 0.05, 8.0)])
def test_compute_stereo_matches_matches_reference(index, mb, mbf):
    (kl, dl, lp, sc, isc), (kr, dr, rp, _, _) = _stereo_case(index)
    ur_o, dep_o, kept, skipped = O.compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, mb, mbf)
    ur_r, dep_r, cnt = R.ref_compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, mb, mbf)
    assert skipped == 0                       # extractor keypoints keep every patch inside the image
    assert kept == cnt and cnt > 50
    assert ur_o.tobytes() == ur_r.tobytes() and dep_o.tobytes() == dep_r.tobytes()


SGOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_stereo_matches.npz")


def test_stereo_oracle_reproduces_reference_golden_vectors():
    """Vectors from the reference's own extractor + ComputeStereoMatches (tools/gen_golden.py)."""
    g = np.load(SGOLDEN)
    assert int(g["count"]) >= 3
    for i in range(int(g["count"])):
        idx, w, h, nf = (int(v) for v in g["cfg_%d" % i][:4])
        mb, mbf = float(g["cfg_%d" % i][4]), float(g["cfg_%d" % i][5])
        left = F.synthetic_frame(idx, w, h); right = F.stereo_right_frame(left, idx)
        side = []
        for img in (left, right):
            ex = O.OracleExtractor(nf, 1.2, 8, 20, 7)
            k, d = ex(img)
            side.append((k, d, [ex.level_pixels(l) for l in range(8)], ex.scale_factors, ex.inv_scale_factors))
        (kl, dl, lp, sc, isc), (kr, dr, rp, _, _) = side
        assert np.array_equal(_sha(kl), g["kl_sha_%d" % i]) and np.array_equal(_sha(kr), g["kr_sha_%d" % i])
        ur, dep, kept, skipped = O.compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, mb, mbf)
        assert skipped == 0 and kept == int(g["n_%d" % i])
        assert ur.tobytes() == g["ur_%d" % i].tobytes() and dep.tobytes() == g["depth_%d" % i].tobytes()


@needs_refm
def test_compute_stereo_matches_random_keypoints_match_reference():
    """Keypoints that are not the extractor's: random positions (kept far enough inside every level that the reference's
    rowRange / colRange stay legal), random octaves, descriptor pairs at every distance, disparities on both sides of the
    accepted range -- the reference's own ComputeStereoMatches against the restatement, 25 frames."""
    rng = np.random.default_rng(2024)
    ex = O.OracleExtractor(300, 1.2, 8, 20, 7)
    for trial in range(25):
        left = F.synthetic_frame(500 + trial, 320, 240)
        right = F.stereo_right_frame(left, 500 + trial, max_disparity=30.0, band=24)
        ex(left); lp = [ex.level_pixels(l).copy() for l in range(8)]
        ex(right); rp = [ex.level_pixels(l).copy() for l in range(8)]
        sc, isc = ex.scale_factors, ex.inv_scale_factors
        n = 220
        kl = np.zeros(n, O.KP_DTYPE)
        kl["octave"] = rng.integers(0, 6, n)
        m = 22.0 * sc[kl["octave"]] + 12.0               # margin in level-0 pixels: patch + sweep stay inside the level
        kl["x"] = (m + rng.random(n) * (320 - 2 * m)).astype(np.float32)
        kl["y"] = (m + rng.random(n) * (240 - 2 * m)).astype(np.float32)
        dl = rng.integers(0, 256, (n, 32)).astype(np.uint8)
        kr = kl.copy()
        kr["x"] = np.clip(kl["x"] - rng.uniform(-6, 45, n).astype(np.float32), m, 320 - m).astype(np.float32)
        kr["y"] = (kl["y"] + rng.uniform(-3, 3, n)).astype(np.float32)
        kr["octave"] = np.clip(kl["octave"] + rng.integers(-2, 3, n), 0, 7)
        flips = np.packbits(rng.random((n, 256)) < rng.uniform(0.0, 0.45, (n, 1)), axis=1)
        dr = dl ^ flips
        perm = rng.permutation(n)
        kr, dr = kr[perm], dr[perm]
        mb, mbf = float(rng.uniform(0.05, 0.6)), float(rng.uniform(5, 40))
        ur_o, dep_o, kept, skipped = O.compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, mb, mbf)
        if skipped:                                           # would be undefined in the reference: not comparable
            continue
        ur_r, dep_r, cnt = R.ref_compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, mb, mbf)
        assert kept == cnt
        assert ur_o.tobytes() == ur_r.tobytes() and dep_o.tobytes() == dep_r.tobytes()


@pytest.mark.skipif(not (R.available() and R.malloc_variant_available()), reason="oracle/_ref/libref_orb_malloc.so not built")
def test_tiebreak_report_reference_under_glibc_malloc():
    """SURVEY 7 hard part 1: the reference's DistributeOctTree orders equal-count nodes by heap address
    (S/ORBextractor.cc:694-698).  Under glibc malloc the same reference build differs from the canonical (bump
    allocator) order in a few keypoints per frame, and even from itself when the heap history changes; everything both
    runs keep is identical.  tools/tiebreak_report.py writes the numbers to profiles/r2_tiebreak.json."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import tiebreak_report as T
    frames = [F.synthetic_frame(i) for i in range(3)]
    r = T.compare(frames)
    assert r["frames"] == 3 and r["keypoints_canonical_total"] > 2900
    assert r["fraction_of_keypoints_not_in_both"] < 0.05          # a handful of leaves per frame, never a different image
    for f in r["per_frame"]:
        assert abs(f["keypoints"] - f["keypoints_glibc"]) <= 16
    committed = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_tiebreak.json")
    assert os.path.exists(committed)
