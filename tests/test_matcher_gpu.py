"""GPU parity: ORBmatcher's three in-scope entry points through the C ABI against the CPU oracle
(oracle/orb_matcher_oracle.c), bit-exact on match indices, counts, distances and vbPrevMatched."""
import numpy as np
import pytest

import oracle_lib as O
from weiner_slamit_v2_b200.matcher import Frame, MapPoints, ORBmatcher
from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, init_pair, projection_frame

pytestmark = pytest.mark.gpu


def test_descriptor_distance_matches_oracle():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5000, 32)).astype(np.uint8)
    b = rng.integers(0, 256, (5000, 32)).astype(np.uint8)
    b[:10] = a[:10]; b[10:20] = ~a[10:20]
    m = ORBmatcher()
    d = m.DescriptorDistance(a, b)
    ref = np.array([O.descriptor_distance(a[i], b[i]) for i in range(len(a))])
    assert np.array_equal(d, ref) and d[:10].max() == 0 and d[10:20].min() == 256
    assert m.DescriptorDistance(a[0], b[0]) == ref[0]
    assert np.array_equal(d, np.unpackbits(a ^ b, axis=1).sum(1))       # property: popcount of the XOR


@pytest.mark.parametrize("brute,window,n", [(False, 100, 1000), (True, 1000, 1000), (False, 10, 700), (True, 40, 333)])
def test_search_for_initialization_matches_oracle(brute, window, n):
    items = 6
    pairs = [init_pair(i, n=n, brute_force=brute) for i in range(items)]
    m = ORBmatcher(0.9, True, max_items=items, max_points=n)
    F1 = [Frame(p[0], p[1], 640, 480) for p in pairs]
    F2 = [Frame(p[2], p[3], 640, 480) for p in pairs]
    nm, m12, pm = m.search_for_initialization_batch(F1, F2, [p[4] for p in pairs], window)
    total = 0
    for i, p in enumerate(pairs):
        on, om12, opm = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, window)
        assert nm[i] == on, (i, nm[i], on)
        assert np.array_equal(m12[i], om12)
        assert np.array_equal(pm[i], opm)
        total += on
    assert total > 0


@pytest.mark.parametrize("ratio", [0.3, 0.6, 0.75, 1.0, 1.5])
def test_search_for_initialization_whole_grid_pruning_is_exact_for_every_ratio(ratio):
    """The whole-grid path drops candidates whose half distance reaches farDist = the smallest d with TH_LOW < ratio * d
    (167, 84, 67, 51, 51 here; 0.3 disables the pruning): results must not depend on it.  Descriptors are close to
    each other (few flipped bits, many near candidates per query) so lists fill, steals and ties happen."""
    items, n = 4, 700
    rng = np.random.default_rng(int(ratio * 100))
    pairs = []
    for i in range(items):
        k1, d1, k2, d2, prev = init_pair(200 + i, n=n, brute_force=True)
        # pull a third of the descriptors towards a few prototypes: several candidates below farDist per query
        proto = rng.integers(0, 256, (6, 32)).astype(np.uint8)
        sel = rng.random(n) < 0.33
        for arr in (d1, d2):
            noise = np.packbits(rng.random((n, 256)) < rng.uniform(0.02, 0.25, (n, 1)), axis=1)
            arr[sel] = (proto[rng.integers(0, 6, n)] ^ noise)[sel]
        pairs.append((k1, d1, k2, d2, prev))
    m = ORBmatcher(ratio, True, max_items=items, max_points=n)
    F1 = [Frame(p[0], p[1], 640, 480) for p in pairs]
    F2 = [Frame(p[2], p[3], 640, 480) for p in pairs]
    nm, m12, pm = m.search_for_initialization_batch(F1, F2, [p[4] for p in pairs], 1000)
    total = 0
    for i, p in enumerate(pairs):
        on, om12, opm = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), ratio, True, 1000)
        assert nm[i] == on, (i, nm[i], on)
        assert np.array_equal(m12[i], om12) and np.array_equal(pm[i], opm)
        total += on
    assert total > 0


def test_search_for_initialization_no_orientation_and_second_round():
    p = init_pair(11, n=800)
    m = ORBmatcher(0.9, False)
    F1, F2 = Frame(p[0], p[1], 640, 480), Frame(p[2], p[3], 640, 480)
    prev = p[4].copy()
    n1, m12 = m.SearchForInitialization(F1, F2, prev, None, 100)
    on, om12, opm = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, False, 100)
    assert n1 == on and np.array_equal(m12, om12) and np.array_equal(prev, opm)
    # second call with the updated vbPrevMatched (what MonocularInitialization does frame after frame)
    n2, m12b = m.SearchForInitialization(F1, F2, prev, None, 100)
    on2, om12b, opm2 = O.search_for_initialization(p[0], p[1], p[2], p[3], opm, (0, 0, 640, 480), 0.9, False, 100)
    assert n2 == on2 and np.array_equal(m12b, om12b) and np.array_equal(prev, opm2)


def test_search_for_initialization_empty_and_ragged():
    m = ORBmatcher(0.9, True, max_items=3, max_points=64)
    p = init_pair(3, n=50)
    empty = np.zeros(0, p[0].dtype)
    F1 = [Frame(p[0], p[1], 640, 480), Frame(empty, np.zeros((0, 32), np.uint8), 640, 480), Frame(p[0][:7], p[1][:7], 640, 480)]
    F2 = [Frame(empty, np.zeros((0, 32), np.uint8), 640, 480), Frame(p[2], p[3], 640, 480), Frame(p[2], p[3], 640, 480)]
    prevs = [p[4], np.zeros((0, 2), np.float32), p[4][:7]]
    nm, m12, pm = m.search_for_initialization_batch(F1, F2, prevs, 100)
    assert nm[0] == 0 and nm[1] == 0 and (m12[0] == -1).all()
    on, om12, opm = O.search_for_initialization(p[0][:7], p[1][:7], p[2], p[3], p[4][:7], (0, 0, 640, 480), 0.9, True, 100)
    assert nm[2] == on and np.array_equal(m12[2], om12) and np.array_equal(pm[2], opm)


@pytest.mark.parametrize("th", [1.0, 3.0, 5.0])
def test_search_by_projection_matches_oracle(th):
    items = 4
    data = [projection_frame(i, n_kp=2000, n_mp=3000) for i in range(items)]
    m = ORBmatcher(0.8, True, max_items=items, max_points=3000)
    frames = [Frame(kp, kd, 1280, 720, SCALE_FACTORS_8) for kp, kd, _ in data]
    # some keypoints already hold a foreign map point (with and without observations)
    rng = np.random.default_rng(5)
    for f in frames:
        idx = rng.choice(f.N, 100, replace=False)
        f.mvpMapPoints[idx] = -2
        f.mvpMapPointObs[idx] = rng.integers(0, 3, 100)
    pre = [(f.mvpMapPoints.copy(), f.mvpMapPointObs.copy()) for f in frames]
    mps = [MapPoints(mp["x"], mp["y"], mp["level"], mp["viewcos"], mp["desc"], mp["in_view"], mp["bad"], mp["xr"], mp["obs"])
           for _, _, mp in data]
    nm = m.search_by_projection_batch(frames, mps, th)
    tot = 0
    for i, (kp, kd, mp) in enumerate(data):
        cnt, kpmp = O.search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, th, pre[i][0], pre[i][1])
        assert nm[i] == cnt, (i, nm[i], cnt)
        assert np.array_equal(frames[i].mvpMapPoints, kpmp)
        tot += cnt
    assert tot > 100


def test_search_by_projection_stereo_and_zero_observation_points():
    kp, kd, mp = projection_frame(42, n_kp=500, n_mp=1500, width=640, height=480)
    rng = np.random.default_rng(1)
    ur = np.where(rng.random(500) < 0.5, kp["x"] - rng.uniform(0, 30, 500), -1).astype(np.float32)
    mp["xr"] = (mp["x"] - rng.uniform(0, 30, 1500)).astype(np.float32)
    mp["kuright"] = ur
    f = Frame(kp, kd, 640, 480, SCALE_FACTORS_8, u_right=ur)
    m = ORBmatcher(0.8)
    n = m.SearchByProjection(f, MapPoints(mp["x"], mp["y"], mp["level"], mp["viewcos"], mp["desc"], mp["in_view"], mp["bad"], mp["xr"], mp["obs"]), 3.0)
    cnt, kpmp = O.search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 640, 480), 0.8, 3.0)
    assert n == cnt and np.array_equal(f.mvpMapPoints, kpmp)


def test_distorted_image_bounds_shift_the_grid():
    """With lens distortion Frame::mnMinX.. are not the image rectangle (S/Frame.cc:561-580)."""
    bounds = (-13.7, -9.2, 661.3, 492.8)
    p = init_pair(21, n=600)
    m = ORBmatcher(0.9, True)
    F1, F2 = Frame(p[0], p[1], 640, 480, bounds=bounds), Frame(p[2], p[3], 640, 480, bounds=bounds)
    prev = p[4].copy()
    n, m12 = m.SearchForInitialization(F1, F2, prev, None, 60)
    on, om12, opm = O.search_for_initialization(p[0], p[1], p[2], p[3], p[4], bounds, 0.9, True, 60)
    assert n == on and np.array_equal(m12, om12) and np.array_equal(prev, opm)
    kp, kd, mp = projection_frame(7, n_kp=800, n_mp=2000, width=640, height=480)
    f = Frame(kp, kd, 640, 480, SCALE_FACTORS_8, bounds=bounds)
    cnt = m.SearchByProjection(f, MapPoints(mp["x"], mp["y"], mp["level"], mp["viewcos"], mp["desc"], mp["in_view"], mp["bad"], mp["xr"], mp["obs"]), 5.0)
    ocnt, okp = O.search_by_projection(mp, kp, kd, SCALE_FACTORS_8, bounds, 0.8, 5.0)
    assert cnt == ocnt and np.array_equal(f.mvpMapPoints, okp)


@pytest.mark.parametrize("th,mode,ori", [(15.0, 0, True), (7.0, 0, True), (15.0, 0, False), (15.0, 1, True), (15.0, 2, True), (40.0, 0, True)])
def test_search_by_projection_last_frame_matches_oracle(th, mode, ori):
    """Scope row N2: SearchByProjection(CurrentFrame, LastFrame, th, bMono) for a batch of frame pairs."""
    from weiner_slamit_v2_b200.workloads import motion_frame
    items = 4
    ws = [motion_frame(i) for i in range(items)]
    bounds = (-13.7, -9.2, 661.3, 492.8)
    m = ORBmatcher(0.9, ori, max_items=items, max_points=2000)
    rng = np.random.default_rng(8)
    frames, pre = [], []
    for w in ws:
        ur = np.where(rng.random(len(w["cur"])) < 0.3, w["cur"]["x"] - rng.uniform(0, 20, len(w["cur"])), -1).astype(np.float32)
        f = Frame(w["cur"], w["cdesc"], 640, 480, SCALE_FACTORS_8, u_right=ur, bounds=bounds)
        ii = rng.choice(f.N, 120, replace=False)
        f.mvpMapPoints[ii[:60]] = -2; f.mvpMapPointObs[ii[:60]] = rng.integers(0, 3, 60)
        f.mvpMapPoints[ii[60:]] = rng.integers(0, len(w["has_mp"]), 60)
        frames.append(f); pre.append((f.mvpMapPoints.copy(), f.mvpMapPointObs.copy(), ur))
    nm = m.search_by_projection_last_frame_batch(frames, ws, th, mode, mbf=40.0)
    tot = 0
    for i, w in enumerate(ws):
        cnt, kpmp = O.search_by_projection_last_frame(w, SCALE_FACTORS_8, bounds, th, mode, ori, 40.0, pre[i][0], pre[i][1], pre[i][2])
        assert nm[i] == cnt, (i, nm[i], cnt)
        assert np.array_equal(frames[i].mvpMapPoints, kpmp)
        tot += cnt
    assert tot > 200


@pytest.mark.parametrize("th,orb_dist,ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False), (25.0, 255, True), (10.0, 0, True)])
def test_search_by_projection_keyframe_matches_oracle(th, orb_dist, ori):
    """Scope row N2, relocalisation overload: SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) for a
    batch of (frame, key frame) pairs of different sizes, incl. an empty key frame; predicted levels (logf) included."""
    from weiner_slamit_v2_b200.workloads import relocalisation_frame
    sizes = [(1500, 2000), (1500, 2000), (700, 900), (0, 300), (40, 60)]
    ws = [relocalisation_frame(40 + i, n_kf=a, n_cur=b) for i, (a, b) in enumerate(sizes)]
    bounds = (-13.7, -9.2, 661.3, 492.8)
    m = ORBmatcher(0.9, ori, max_items=len(ws), max_points=2000)
    frames = []
    for w in ws:
        f = Frame(w["cur"], w["cdesc"], 640, 480, SCALE_FACTORS_8, bounds=bounds)
        f.mvpMapPoints[:] = w["kp_mp"]
        frames.append(f)
    nm = m.search_by_projection_keyframe_batch(frames, ws, th, orb_dist)
    tot = 0
    for i, w in enumerate(ws):
        cnt, kpmp = O.search_by_projection_keyframe(w, SCALE_FACTORS_8, bounds, th, orb_dist, ori)
        assert nm[i] == cnt, (i, nm[i], cnt)
        assert np.array_equal(frames[i].mvpMapPoints, kpmp), i
        tot += cnt
    assert tot > 400 or orb_dist == 0


def test_search_by_projection_keyframe_reproduces_reference_golden_vectors():
    """tests/golden/ref_match_keyframe.npz was produced by the reference's own ORBmatcher.cc (tools/gen_golden.py)."""
    import os
    from weiner_slamit_v2_b200.workloads import relocalisation_frame
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_keyframe.npz"))
    bounds = (-13.7, -9.2, 661.3, 492.8)
    for i in range(int(g["count"])):
        c = g["cfg_%d" % i]
        w = relocalisation_frame(int(c[0]))
        m = ORBmatcher(0.9, bool(c[3]), max_items=1, max_points=2000)
        f = Frame(w["cur"], w["cdesc"], 640, 480, SCALE_FACTORS_8, bounds=bounds)
        f.mvpMapPoints[:] = w["kp_mp"]
        nm = m.search_by_projection_keyframe_batch([f], [w], float(c[1]), int(c[2]))
        assert nm[0] == int(g["n_%d" % i]) and np.array_equal(f.mvpMapPoints, g["kpmp_%d" % i])


@pytest.mark.parametrize("ratio,ori", [(0.7, True), (0.9, True), (0.7, False)])
def test_search_by_bow_matches_oracle(ratio, ori):
    """Scope row N3: SearchByBoW(pKF, F, vpMapPointMatches) for a ragged batch: ~20 features per node, nodes with
    several hundred features (lists longer than a warp), a single node holding everything, empty sides."""
    from weiner_slamit_v2_b200.workloads import bow_pair
    cfg = [(2000, 2000, 100), (2000, 1500, 100), (500, 800, 30), (0, 100, 10), (100, 0, 10), (50, 50, 1), (1000, 1000, 1000),
           (900, 1100, 2), (1500, 1500, 1)]
    ws = [bow_pair(70 + i, *c) for i, c in enumerate(cfg)]
    m = ORBmatcher(ratio, ori, max_items=len(ws), max_points=2000)
    nm, matches = m.search_by_bow_batch(ws)
    tot = 0
    for i, w in enumerate(ws):
        cnt, mo = O.search_by_bow(w, ratio, ori)
        assert nm[i] == cnt, (i, nm[i], cnt)
        assert np.array_equal(matches[i], mo), i
        tot += cnt
    assert tot > 1500
    # SearchByBoW(pKF1, pKF2, vpMatches12) (S/ORBmatcher.cc:526-659) on the same pairs
    nm, matches = m.search_by_bow_batch(ws, keyframes=True)
    tot = 0
    for i, w in enumerate(ws):
        cnt, mo = O.search_by_bow_keyframes(w, ratio, ori)
        assert nm[i] == cnt, (i, nm[i], cnt)
        assert np.array_equal(matches[i], mo), i
        tot += cnt
    assert tot > 1000


def test_search_by_bow_reproduces_reference_golden_vectors():
    """tests/golden/ref_match_bow.npz was produced by the reference's own ORBmatcher.cc (tools/gen_golden.py)."""
    import os
    from weiner_slamit_v2_b200.workloads import bow_pair
    for name, kk in (("ref_match_bow.npz", False), ("ref_match_bowkf.npz", True)):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            m = ORBmatcher(float(c[4]), bool(c[5]), max_items=1, max_points=2000)
            nm, matches = m.search_by_bow_batch([bow_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3]))], keyframes=kk)
            assert nm[0] == int(g["n_%d" % i]) and np.array_equal(matches[0], g["m_%d" % i])


@pytest.mark.parametrize("only_stereo,ori", [(False, True), (True, True), (False, False)])
def test_search_for_triangulation_matches_oracle(only_stereo, ori):
    """Scope row N3: SearchForTriangulation for a ragged batch of key-frame pairs (sideways / forward motion, mono /
    partly stereo, a single node holding everything, empty sides), and the reference's golden vectors."""
    from weiner_slamit_v2_b200.workloads import triangulation_pair
    cfg = [(2000, 2000, 100, 0.0, False), (2000, 1500, 100, 0.3, False), (2000, 2000, 100, 0.0, True), (1200, 1200, 50, 0.3, True),
           (0, 100, 10, 0.0, False), (100, 0, 10, 0.0, False), (50, 50, 1, 0.5, False), (1000, 1000, 1000, 0.0, True), (1500, 1500, 1, 0.0, True)]
    ws = [triangulation_pair(90 + i, a, b, nn, stereo_fraction=sfr, forward=fwd) for i, (a, b, nn, sfr, fwd) in enumerate(cfg)]
    m = ORBmatcher(0.6, ori, max_items=len(ws), max_points=2000)
    nm, matches = m.search_for_triangulation_batch(ws, only_stereo)
    tot = 0
    for i, w in enumerate(ws):
        cnt, mo = O.search_for_triangulation(w, only_stereo, ori)
        assert nm[i] == cnt, (i, nm[i], cnt)
        assert np.array_equal(matches[i], mo), i
        tot += cnt
    assert tot > (20 if only_stereo else 400)
    if not only_stereo and ori:
        import os
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_triangulation.npz"))
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            w = triangulation_pair(int(c[0]), int(c[1]), int(c[2]), int(c[3]), stereo_fraction=float(c[4]), forward=bool(c[5]))
            mm = ORBmatcher(0.6, bool(c[7]), max_items=1, max_points=2000)
            nm1, m1 = mm.search_for_triangulation_batch([w], bool(c[6]))
            assert nm1[0] == int(g["n_%d" % i]) and np.array_equal(m1[0], g["m_%d" % i])


def test_distinctive_descriptors_match_oracle():
    """Scope row N4: the selection of MapPoint::ComputeDistinctiveDescriptors for a ragged batch of map points:
    0, 1, 2 observations, exactly a warp, more than a warp, several hundred."""
    from weiner_slamit_v2_b200.workloads import observed_descriptors
    sizes = [0, 1, 2, 3, 4, 5, 7, 8, 16, 31, 32, 33, 50, 64, 100, 257] + [int(v) for v in np.random.default_rng(0).integers(1, 40, 300)]
    obs = observed_descriptors(1, sizes)
    m = ORBmatcher(0.9, True, max_items=1, max_points=16)
    best, med = m.distinctive_descriptors_batch(obs)
    for p, d in enumerate(obs):
        idx, mo = O.distinctive_descriptor(d)
        assert best[p] == idx, (p, len(d), best[p], idx)
        assert idx < 0 or med[p] == mo


@pytest.mark.parametrize("th", [2.5, 3.0, 10.0])
def test_fuse_search_matches_oracle(th):
    """Scope row N3: the search of Fuse(pKF, vpMapPoints, th) for a ragged batch, and the reference's golden vectors."""
    from weiner_slamit_v2_b200.workloads import fuse_frame
    bounds = (-13.7, -9.2, 661.3, 492.8)
    ws = [fuse_frame(50 + i, a, b, stereo_fraction=s) for i, (a, b, s) in enumerate([(3000, 2000, 0.2), (3000, 2000, 0.0), (500, 300, 0.5),
                                                                                       (0, 100, 0.2), (100, 0, 0.2)])]
    m = ORBmatcher(0.6, True, max_items=len(ws), max_points=3000)
    res = m.fuse_search_batch(ws, bounds, th)
    tot = 0
    for i, w in enumerate(ws):
        bo, do = O.fuse_search(w, bounds, th)
        assert np.array_equal(res[i][0], bo), i
        assert np.array_equal(res[i][1], do), i
        tot += int((bo >= 0).sum())
    assert tot > 400
    if th == 3.0:
        import os
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_fuse.npz"))
        for i in range(int(g["count"])):
            c = g["cfg_%d" % i]
            r = m.fuse_search_batch([fuse_frame(int(c[0]), int(c[1]), int(c[2]))], bounds, float(c[3]))
            assert np.array_equal(r[0][0], g["best_%d" % i])


def test_fuse_sim3_search_and_search_by_sim3_match_oracle():
    """Scope row N3: the search of Fuse(pKF, Scw, ...) (mode 1) and SearchBySim3 (two mode-2 legs + agreement), ragged
    batches, plus the reference's golden vectors."""
    import os
    from weiner_slamit_v2_b200.workloads import fuse_frame, sim3_pair
    bounds = (-13.7, -9.2, 661.3, 492.8)
    ws = [fuse_frame(60 + i, a, b) for i, (a, b) in enumerate([(3000, 2000), (2000, 1500), (500, 300), (0, 100), (100, 0)])]
    m = ORBmatcher(0.6, True, max_items=len(ws), max_points=3000)
    res = m.fuse_search_batch(ws, bounds, 4.0, mode=1)
    for i, w in enumerate(ws):
        bo, do = O.fuse_search(w, bounds, 4.0, 1)
        assert np.array_equal(res[i][0], bo) and np.array_equal(res[i][1], do), i
    ss = [sim3_pair(60 + i, a, b) for i, (a, b) in enumerate([(1500, 1500), (1500, 1200), (300, 400), (0, 100), (100, 0)])]
    found, matches = m.search_by_sim3_batch(ss, bounds, 7.5)
    tot = 0
    for i, w in enumerate(ss):
        fo, mo = O.search_by_sim3(w, bounds, 7.5)
        assert found[i] == fo and np.array_equal(matches[i], mo), i
        tot += fo
    assert tot > 150
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_sim3.npz"))
    for i in range(int(g["fcount"])):
        c = g["fcfg_%d" % i]
        r = m.fuse_search_batch([fuse_frame(int(c[0]), int(c[1]), int(c[2]))], bounds, float(c[3]), mode=1)
        assert np.array_equal(r[0][0], g["fbest_%d" % i])
    for i in range(int(g["scount"])):
        c = g["scfg_%d" % i]
        f, mm = m.search_by_sim3_batch([sim3_pair(int(c[0]), int(c[1]), int(c[2]))], bounds, float(c[3]))
        assert f[0] == int(g["sn_%d" % i]) and np.array_equal(mm[0], g["sm_%d" % i])


@pytest.mark.parametrize("th", [4, 10])
def test_search_by_projection_sim3_matches_oracle(th):
    """Scope row N3: SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) for a ragged batch with pre-occupied keypoints,
    and the reference's golden vectors."""
    import os
    from weiner_slamit_v2_b200.workloads import fuse_frame
    bounds = (-13.7, -9.2, 661.3, 492.8)
    rng = np.random.default_rng(6)
    ws = [fuse_frame(80 + i, a, b) for i, (a, b) in enumerate([(3000, 2000), (2000, 1500), (500, 300), (0, 100), (100, 0), (6000, 1000)])]
    pre = [np.where(rng.random(len(w["kp"])) < 0.1, -2, -1).astype(np.int32) for w in ws]
    m = ORBmatcher(0.75, True, max_items=len(ws), max_points=6000)
    cnt, matched = m.search_by_projection_sim3_batch(ws, bounds, th, pre)
    tot = 0
    for i, w in enumerate(ws):
        co, mo = O.search_by_projection_sim3(w, bounds, th, pre[i])
        assert cnt[i] == co and np.array_equal(matched[i], mo), i
        tot += co
    assert tot > 800
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_match_projsim3.npz"))
    for i in range(int(g["count"])):
        c = g["cfg_%d" % i]
        if int(c[3]) != th:
            continue
        p0 = np.where(np.random.default_rng(int(c[0])).random(int(c[2])) < 0.1, -2, -1).astype(np.int32)
        cc, mm = m.search_by_projection_sim3_batch([fuse_frame(int(c[0]), int(c[1]), int(c[2]))], bounds, th, [p0])
        assert cc[0] == int(g["n_%d" % i]) and np.array_equal(mm[0], g["m_%d" % i])


def test_bow_transform_matches_oracle_and_reference_golden_vectors():
    """Scope row N4: Frame::ComputeBoW (DBoW2 transform) for a ragged batch of frames on several vocabularies: words,
    bit-identical L1-normalised tf-idf doubles, FeatureVector; then the vectors the reference's own DBoW2 produced."""
    import os
    from weiner_slamit_v2_b200.matcher import Vocabulary, bow_transform_batch
    from weiner_slamit_v2_b200.workloads import synthetic_vocabulary, vocabulary_features
    m = ORBmatcher(0.7, True, max_items=6, max_points=2048)

    def same(a, b):
        return all(np.array_equal(a[x], b[x]) for x in ("word", "node", "start", "feat")) and np.asarray(a["value"]).tobytes() == np.asarray(b["value"]).tobytes()
    for vi, (k, L) in enumerate([(6, 4), (10, 3), (3, 6), (4, 2)]):
        voc = synthetic_vocabulary(30 + vi, k, L)
        V = Vocabulary(voc)
        for lu in (4, 1, 10):
            descs = [vocabulary_features(vi * 10 + j, voc, n) for j, n in enumerate([2000, 500, 1, 0, 1000, 33])]
            res = bow_transform_batch(m, V, descs, lu)
            for j, d in enumerate(descs):
                assert same(res[j], O.bow_transform(voc, d, lu)), (vi, lu, j)
        V.close()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_bow_transform.npz"))
    for i in range(int(g["count"])):
        vi, k, L, n, lu = (int(v) for v in g["cfg_%d" % i])
        voc = synthetic_vocabulary(vi, k, L)
        V = Vocabulary(voc)
        r = bow_transform_batch(m, V, [vocabulary_features(vi, voc, n)], lu)[0]
        assert same(r, {key: g["%s_%d" % (key, i)] for key in ("word", "value", "node", "start", "feat")})
        V.close()
