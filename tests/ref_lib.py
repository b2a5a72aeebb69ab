"""ctypes binding of oracle/_ref/libref_orb.so: the reference's OWN unmodified ORBextractor.cc
compiled for x86 against mini-cv (see oracle/ref_harness.cc).  TEST INFRASTRUCTURE ONLY."""
import contextlib
import ctypes as C
import os

import numpy as np

from oracle_lib import KP_DTYPE, ORACLE_DIR, build_oracle

REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_orb.so")
_u8p = C.POINTER(C.c_uint8)
_lib = None


def available():
    return os.path.exists(REF_SO)


def _load_extractor_library(path):
    L = C.CDLL(path)
    L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ref_extractor_create.restype = C.c_void_p
    L.ref_extractor_destroy.argtypes = [C.c_void_p]
    L.ref_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.ref_extract.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, C.c_int]
    L.ref_extract.restype = C.c_int
    L.ref_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ref_level_pixels.argtypes = [C.c_void_p, C.c_int, C.c_int, _u8p]
    L.ref_distribute_octree.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.ref_distribute_octree.restype = C.c_int
    L.ref_extract_batch_mt.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.ref_set_alloc_direction.argtypes = [C.c_int]
    L.minicv_set_blur_variant.argtypes = [C.c_int]
    return L


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = _load_extractor_library(REF_SO)
    return _lib


# the same reference build WITHOUT the bump allocator (glibc malloc decides the quadtree's pointer-valued tie-break,
# S/ORBextractor.cc:694-698): what an integrator who links the real reference sees.  tools/tiebreak_report.py
REF_MALLOC_SO = os.path.join(ORACLE_DIR, "_ref", "libref_orb_malloc.so")
_malloc_lib = None


def malloc_variant_available():
    return os.path.exists(REF_MALLOC_SO)


def malloc_lib():
    global _malloc_lib
    if _malloc_lib is None:
        build_oracle()
        _malloc_lib = _load_extractor_library(REF_MALLOC_SO)
    return _malloc_lib


class RefExtractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, default_malloc=False):
        self.L = malloc_lib() if default_malloc else lib()
        self.h = self.L.ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32); um = np.zeros(16, np.int32); pat = np.zeros(1024, np.int32)
        self.L.ref_tables(self.h, *(a.ctypes.data for a in (sc, isc, s2, is2, per, um, pat)))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per, umax=um, pattern=pat)

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 4 * 64
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.L.ref_extract(self.h, img.ctypes.data_as(_u8p), w, h, img.strides[0], kps.ctypes.data,
                               desc.ctypes.data_as(_u8p), cap)
        assert n >= 0
        return kps[:n].copy(), desc[:n].copy()

    def level_pixels(self, level, with_border=False):
        w = C.c_int(); h = C.c_int()
        self.L.ref_level_size(self.h, level, C.byref(w), C.byref(h))
        b = 19 if with_border else 0
        out = np.zeros((h.value + 2 * b, w.value + 2 * b), np.uint8)
        self.L.ref_level_pixels(self.h, level, int(with_border), out.ctypes.data_as(_u8p))
        return out

    def distribute_octree(self, cands, minX, maxX, minY, maxY, N, level=0):
        cands = np.ascontiguousarray(cands, KP_DTYPE)
        cap = N + 4 * 64 + 16
        out = np.zeros(cap, KP_DTYPE)
        n = self.L.ref_distribute_octree(self.h, cands.ctypes.data, len(cands), minX, maxX, minY, maxY, N, level,
                                         out.ctypes.data, cap)
        assert n >= 0
        return out[:n]


def extract_batch_mt(frames, threads, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    counts = np.zeros(n, np.int32)
    lib().ref_extract_batch_mt(nfeatures, scale_factor, nlevels, ini_th, min_th, frames.ctypes.data_as(_u8p),
                               n, w, h, threads, counts.ctypes.data_as(C.POINTER(C.c_int)))
    return counts


# ---- the reference's own ORBmatcher.cc (+ Frame.cc, MapPoint.cc): oracle/_ref/libref_matcher.so -------------
REFM_SO = os.path.join(ORACLE_DIR, "_ref", "libref_matcher.so")
_mlib = None
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def matcher_available():
    return os.path.exists(REFM_SO)


def _load_matcher_library(path):
    L = C.CDLL(path)
    L.refm_descriptor_distance.argtypes = [_u8p, _u8p]
    L.refm_search_for_initialization.argtypes = ([C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p] * 2 +
                                                 [_f32p, C.c_float, C.c_int, C.c_int, _f32p, _i32p])
    L.refm_search_by_projection.argtypes = [
        C.c_int, _u8p, _u8p, _f32p, _f32p, _f32p, _i32p, _f32p, _u8p, _i32p,
        C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p, _i32p, _i32p,
        C.c_int, _f32p, _f32p, C.c_float, C.c_float]
    return L


def mlib():
    global _mlib
    if _mlib is None:
        _mlib = _load_matcher_library(REFM_SO)
    return _mlib


# ---- the same harness with the DROP-IN bodies linked in place of the reference's (oracle/Makefile, `make shimm`):
# every ref_* wrapper below then drives weiner_slamit_v2_b200/shim/*.cc on real ORB_SLAM2::Frame / MapPoint / KeyFrame
# objects, and the search itself runs on the GPU through liborb_b200.so.
SHIMM_SO = os.path.join(ORACLE_DIR, "_ref", "libshim_matcher.so")
_shimlib = None


def shim_available():
    return os.path.exists(SHIMM_SO)


def shimlib():
    global _shimlib
    if _shimlib is None:
        _shimlib = _load_matcher_library(SHIMM_SO)
    return _shimlib


@contextlib.contextmanager
def shim_bodies():
    """Inside this block the ref_* wrappers call libshim_matcher.so instead of libref_matcher.so."""
    global _mlib
    saved = _mlib
    _mlib = shimlib()
    try:
        yield _mlib
    finally:
        _mlib = saved


def _p(a, t):
    return a.ctypes.data_as(t)


def _f(a): return np.ascontiguousarray(a, np.float32)
def _i(a): return np.ascontiguousarray(a, np.int32)
def _b(a): return np.ascontiguousarray(a, np.uint8)


def ref_descriptor_distance(a, b):
    a, b = _b(a), _b(b)
    return mlib().refm_descriptor_distance(_p(a, _u8p), _p(b, _u8p))


def ref_search_for_initialization(k1, d1, k2, d2, prev_matched, bounds, nnratio=0.9, check_ori=True, window=100):
    n1, n2 = len(k1), len(k2)
    a = [_f(k1["x"]), _f(k1["y"]), _i(k1["octave"]), _f(k1["angle"]), _b(d1) if n1 else np.zeros((1, 32), np.uint8)]
    b = [_f(k2["x"]), _f(k2["y"]), _i(k2["octave"]), _f(k2["angle"]), _b(d2) if n2 else np.zeros((1, 32), np.uint8)]
    bnd = _f(bounds).reshape(4)
    pm = _f(prev_matched).copy().reshape(-1, 2) if n1 else np.zeros((1, 2), np.float32)
    m12 = np.full(max(n1, 1), -1, np.int32)
    n = mlib().refm_search_for_initialization(
        n1, _p(a[0], _f32p), _p(a[1], _f32p), _p(a[2], _i32p), _p(a[3], _f32p), _p(a[4], _u8p),
        n2, _p(b[0], _f32p), _p(b[1], _f32p), _p(b[2], _i32p), _p(b[3], _f32p), _p(b[4], _u8p),
        _p(bnd, _f32p), nnratio, int(check_ori), window, _p(pm, _f32p), _p(m12, _i32p))
    return n, m12[:n1], pm[:n1]


def ref_search_by_projection(mp, kp, kdesc, scale_factors, bounds, nnratio=0.8, th=1.0, kp_mp=None, kp_mp_obs=None):
    nmp, n = len(mp["x"]), len(kp)
    kx, ky, ko = _f(kp["x"]), _f(kp["y"]), _i(kp["octave"])
    kur = _f(mp.get("kuright", np.full(n, -1.0, np.float32)))
    kdesc = _b(kdesc)
    kp_mp = np.full(max(n, 1), -1, np.int32) if kp_mp is None else _i(kp_mp).copy()
    kp_mp_obs = np.zeros(max(n, 1), np.int32) if kp_mp_obs is None else _i(kp_mp_obs)
    iv, bad = _b(mp["in_view"]), _b(mp["bad"])
    x, y, xr, lv, vc = _f(mp["x"]), _f(mp["y"]), _f(mp["xr"]), _i(mp["level"]), _f(mp["viewcos"])
    de, ob = _b(mp["desc"]), _i(mp["obs"])
    sf, bnd = _f(scale_factors), _f(bounds).reshape(4)
    cnt = mlib().refm_search_by_projection(
        nmp, _p(iv, _u8p), _p(bad, _u8p), _p(x, _f32p), _p(y, _f32p), _p(xr, _f32p), _p(lv, _i32p), _p(vc, _f32p),
        _p(de, _u8p), _p(ob, _i32p), n, _p(kx, _f32p), _p(ky, _f32p), _p(ko, _i32p), _p(kur, _f32p), _p(kdesc, _u8p),
        _p(kp_mp, _i32p), _p(kp_mp_obs, _i32p), len(sf), _p(sf, _f32p), _p(bnd, _f32p), nnratio, th)
    return cnt, kp_mp[:n]


def ref_search_by_projection_last_frame(w, scale_factors, bounds, th=15.0, check_ori=True, mbf=40.0, kp_mp=None,
                                        kp_mp_obs=None, kuright=None):
    """The reference's own SearchByProjection(CurrentFrame, LastFrame, th, bMono=true)."""
    import oracle_lib as O
    L = mlib()
    L.refm_search_by_projection_last_frame.argtypes = [
        C.c_int, _u8p, _u8p, _f32p, _u8p, _i32p, _i32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
        C.c_int, _f32p, _f32p, _i32p, _f32p, _f32p, _u8p, _i32p, _i32p,
        C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_int]
    L.refm_search_by_projection_last_frame.restype = C.c_int
    # same argument order as the oracle; the reference takes bMono where the oracle takes mode (mono <-> mode 0)
    def call(*args):
        args = list(args)
        args[-2] = 1
        return L.refm_search_by_projection_last_frame(*args)
    return O.search_by_projection_last_frame(w, scale_factors, bounds, th, 0, check_ori, mbf, kp_mp, kp_mp_obs, kuright, fn=call)


def ref_search_by_projection_keyframe(w, scale_factors, bounds, th=10.0, orb_dist=100, check_ori=True, kp_mp=None):
    """The reference's own SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist).  Also returns Ow as
    the harness's cv::Mat expression evaluates it."""
    import oracle_lib as O
    L = mlib()
    L.refm_search_by_projection_keyframe.argtypes = [
        C.c_int, _u8p, _f32p, _u8p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
        C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p, _i32p, C.c_int, _f32p, C.c_float, _f32p,
        C.c_float, C.c_int, C.c_int, _f32p]
    L.refm_search_by_projection_keyframe.restype = C.c_int
    a, n = O._keyframe_args(w, kp_mp)
    sf, bnd = O._f(scale_factors), O._bounds(bounds)
    ow = np.zeros(3, np.float32)
    P = O._ptr
    cnt = L.refm_search_by_projection_keyframe(
        len(a["valid"]), P(a["valid"], _u8p), P(a["wpos"], _f32p), P(a["mp_desc"], _u8p), P(a["mfmax"], _f32p),
        P(a["mfmin"], _f32p), P(a["kfang"], _f32p), P(a["R"], _f32p), P(a["t"], _f32p), P(a["K"], _f32p),
        n, P(a["kx"], _f32p), P(a["ky"], _f32p), P(a["ko"], _i32p), P(a["ka"], _f32p), P(a["kd"], _u8p),
        P(a["kpmp"], _i32p), len(sf), P(sf, _f32p), float(w["log_scale"]), P(bnd, _f32p), th, orb_dist, int(check_ori),
        P(ow, _f32p))
    return cnt, a["kpmp"][:n], ow


def ref_search_by_bow(w, nnratio=0.7, check_ori=True):
    """The reference's own SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)."""
    import oracle_lib as O
    return O._bow_call(mlib().refm_search_by_bow, w, nnratio, check_ori, lambda v: v)


def ref_search_by_bow_keyframes(w, nnratio=0.75, check_ori=True):
    """The reference's own SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)."""
    import oracle_lib as O
    return O._bow_kk_call(mlib().refm_search_by_bow_keyframes, w, nnratio, check_ori, lambda v: v)


def ref_search_for_triangulation(w, only_stereo=False, check_ori=True):
    """The reference's own SearchForTriangulation; also returns the epipole its expressions give."""
    import oracle_lib as O
    _u32p = O._u32p
    P = O._ptr
    f = mlib().refm_search_for_triangulation
    f.argtypes = [C.c_int, _u8p, _u8p, _f32p, _f32p, _f32p, _f32p, C.c_int, _u32p, _i32p, _u32p,
                  C.c_int, _u8p, _u8p, _f32p, _f32p, _i32p, _f32p, _f32p, C.c_int, _u32p, _i32p, _u32p,
                  _f32p, _f32p, _f32p, _f32p, C.c_int, _f32p, _f32p, C.c_int, C.c_int, _i32p, _f32p]
    f.restype = C.c_int
    A, B = O._tri_sides(w)
    n1 = len(A["x"])
    m = np.full(max(n1, 1), -1, np.int32)
    ep = np.zeros(2, np.float32)
    F, sf, ls = O._f(w["F12"]), O._f(w["scale_factors"]), O._f(w["level_sigma2"])
    pose2 = O._f(np.concatenate([w["R2w"], w["t2w"]]))
    cw, K = O._f(w["Cw1"]), O._f(w["K"])
    cnt = f(n1, P(A["mp"], _u8p), P(A["d"], _u8p), P(A["x"], _f32p), P(A["y"], _f32p), P(A["an"], _f32p), P(A["ur"], _f32p),
            len(A["nd"]), P(A["nd"], _u32p), P(A["st"], _i32p), P(A["ft"], _u32p),
            len(B["x"]), P(B["mp"], _u8p), P(B["d"], _u8p), P(B["x"], _f32p), P(B["y"], _f32p), P(B["oc"], _i32p),
            P(B["an"], _f32p), P(B["ur"], _f32p), len(B["nd"]), P(B["nd"], _u32p), P(B["st"], _i32p), P(B["ft"], _u32p),
            P(F, _f32p), P(cw, _f32p), P(pose2, _f32p), P(K, _f32p), len(sf), P(sf, _f32p), P(ls, _f32p),
            int(only_stereo), int(check_ori), P(m, _i32p), P(ep, _f32p))
    return cnt, m[:n1], ep


def ref_compute_distinctive_descriptors(desc, bad=None):
    """The reference's own MapPoint::ComputeDistinctiveDescriptors.  Returns the chosen 32-byte descriptor or None."""
    import oracle_lib as O
    L = mlib()
    L.refm_compute_distinctive_descriptors.argtypes = [C.c_int, _u8p, _u8p, _u8p]
    L.refm_compute_distinctive_descriptors.restype = C.c_int
    d = O._b(desc).reshape(-1, 32)
    b = O._b(np.zeros(len(d), np.uint8) if bad is None else bad)
    out = np.zeros(32, np.uint8)
    ok = L.refm_compute_distinctive_descriptors(len(d), O._ptr(d, _u8p), O._ptr(b, _u8p), O._ptr(out, _u8p))
    return out if ok else None


def ref_fuse_search(w, bounds, th=3.0):
    """The reference's own ORBmatcher::Fuse(pKF, {pMP}, th), one candidate map point per call (see the harness)."""
    import oracle_lib as O
    return O._fuse_call(mlib().refm_fuse_search, w, bounds, th, lambda v: v, False)[0]


def ref_fuse_search_sim3(w, bounds, th=3.0):
    """The reference's own Fuse(pKF, Scw, vpPoints, th, vpReplacePoint), one candidate per call, Scw = [Rcw | tcw]."""
    import oracle_lib as O
    return O._fuse_call(mlib().refm_fuse_search_sim3, w, bounds, th, lambda v: v, False)[0]


def ref_search_by_sim3(w, bounds, th=7.5):
    """The reference's own SearchBySim3 (s12 = 1, no previous matches)."""
    import oracle_lib as O
    P = O._ptr
    f = mlib().refm_search_by_sim3
    side = [C.c_int, _u8p, _f32p, _u8p, _f32p, _f32p, _f32p, _f32p, _i32p, _u8p, _f32p, _f32p]
    f.argtypes = side + side + [_f32p, C.c_int, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, C.c_float, _i32p]
    f.restype = C.c_int
    keep = []

    def args(k):
        a = [O._b(k["valid"]), O._f(k["wpos"]), O._b(k["mp_desc"]), O._f(k["mf_max"]), O._f(k["mf_min"]), O._f(k["kp"]["x"]), O._f(k["kp"]["y"]),
             O._i(k["kp"]["octave"]), O._b(k["kdesc"]), O._f(k["Rcw"]), O._f(k["tcw"])]
        keep.extend(a)
        return [len(a[0])] + [P(x, t) for x, t in zip(a, side[1:])]
    c = w["common"]
    sf, il, b = O._f(c["scale_factors"]), O._f(c["inv_level_sigma2"]), O._bounds(bounds)
    K, R12, t12 = O._f(c["K"]), O._f(w["R12"]), O._f(w["t12"])
    n1 = len(w["k1"]["valid"])
    m = np.full(max(n1, 1), -1, np.int32)
    found = f(*(args(w["k1"]) + args(w["k2"]) + [P(K, _f32p), len(sf), P(sf, _f32p), P(il, _f32p), float(c["log_scale"]), P(b, _f32p),
                                                   P(R12, _f32p), P(t12, _f32p), float(th), P(m, _i32p)]))
    return found, m[:n1]


def ref_search_by_projection_sim3(w, bounds, th=10, matched=None):
    """The reference's own SearchByProjection(pKF, Scw, vpPoints, vpMatched, th), Scw = [Rcw | tcw]."""
    import oracle_lib as O
    return O._proj_sim3_call(mlib().refm_search_by_projection_sim3, w, bounds, th, lambda v: v, matched)


def ref_bow_transform(voc_path, desc, levelsup=4):
    """The reference's own DBoW2 transform on the vocabulary its loadFromTextFile reads from voc_path."""
    import oracle_lib as O
    f = mlib().refm_bow_transform
    f.argtypes = [C.c_char_p, C.c_int, _u8p, C.c_int, C.POINTER(C.c_int32), O._u32p, O._f64p, C.POINTER(C.c_int32), O._u32p, _i32p, O._u32p]
    f.restype = C.c_int
    d = O._b(desc).reshape(-1, 32)
    o = O._bow_out(len(d))
    rc = f(voc_path.encode(), len(d), O._ptr(d, _u8p), int(levelsup), C.byref(o["bn"]), O._ptr(o["bw"], O._u32p), O._ptr(o["bv"], O._f64p),
           C.byref(o["fn"]), O._ptr(o["fnode"], O._u32p), O._ptr(o["fstart"], _i32p), O._ptr(o["ffeat"], O._u32p))
    assert rc == 0, "the reference could not load the vocabulary"
    return O._bow_result(o)


def ref_compute_stereo_matches(kl, dl, kr, dr, scale, inv_scale, lpyr, rpyr, mb, mbf):
    """The reference's own Frame::ComputeStereoMatches (S/Frame.cc:591-763).  Returns (u_right, depth, count)."""
    import oracle_lib as O
    f = mlib().refm_compute_stereo_matches
    f.argtypes = [C.c_int, _f32p, _f32p, _i32p, _u8p, C.c_int, _f32p, _f32p, _i32p, _u8p, C.c_int, _f32p, _f32p,
                  C.POINTER(_u8p), _i32p, C.POINTER(_u8p), _i32p, _i32p, _i32p, C.c_float, C.c_float, _f32p, _f32p]
    f.restype = C.c_int
    n, nr = len(kl), len(kr)
    a = [_f(kl["x"]), _f(kl["y"]), _i(kl["octave"]), _b(dl) if n else np.zeros((1, 32), np.uint8)]
    b = [_f(kr["x"]), _f(kr["y"]), _i(kr["octave"]), _b(dr) if nr else np.zeros((1, 32), np.uint8)]
    sc, isc = _f(scale), _f(inv_scale)
    lk, lp, lpitch = O._pyr_args(lpyr)
    rk, rp, rpitch = O._pyr_args(rpyr)
    lw = np.array([p.shape[1] for p in lk], np.int32); lh = np.array([p.shape[0] for p in lk], np.int32)
    ur = np.zeros(max(n, 1), np.float32); dep = np.zeros(max(n, 1), np.float32)
    cnt = f(n, _p(a[0], _f32p), _p(a[1], _f32p), _p(a[2], _i32p), _p(a[3], _u8p),
            nr, _p(b[0], _f32p), _p(b[1], _f32p), _p(b[2], _i32p), _p(b[3], _u8p), len(sc), _p(sc, _f32p), _p(isc, _f32p),
            lp, _p(lpitch, _i32p), rp, _p(rpitch, _i32p), _p(lw, _i32p), _p(lh, _i32p), mb, mbf, _p(ur, _f32p), _p(dep, _f32p))
    return ur[:n], dep[:n], cnt
