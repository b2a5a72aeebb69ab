"""ctypes binding of the CPU oracle (oracle/liborb_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def _ptr(a, t):
    return a.ctypes.data_as(t)


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liborb_oracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("orb_oracle.c", "orb_matcher_oracle.c", "orb_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.orc_round.argtypes = [C.c_float]; L.orc_round.restype = C.c_int
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]; L.orc_fast_atan2.restype = C.c_float
        L.orc_sincosf.argtypes = [C.c_float, _f32p, _f32p]
        L.orc_resize_linear_u8.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
        L.orc_copy_make_border_reflect101.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
        L.orc_fast9_16.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_fast9_16.restype = C.c_int
        L.orc_gaussian_blur7.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, C.c_int, C.c_int]
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_create.restype = C.c_void_p
        L.orc_extractor_destroy.argtypes = [C.c_void_p]
        L.orc_extractor_set_blur_variant.argtypes = [C.c_void_p, C.c_int]
        for name in ("orc_scale_factors", "orc_inv_scale_factors", "orc_level_sigma2", "orc_inv_level_sigma2"):
            getattr(L, name).argtypes = [C.c_void_p]; getattr(L, name).restype = _f32p
        for name in ("orc_features_per_level", "orc_umax"):
            getattr(L, name).argtypes = [C.c_void_p]; getattr(L, name).restype = C.POINTER(C.c_int)
        L.orc_pattern.restype = C.POINTER(C.c_int8)
        L.orc_extract.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, C.c_int]
        L.orc_extract.restype = C.c_int
        for name in ("orc_level_width", "orc_level_height"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]; getattr(L, name).restype = C.c_int
        for name in ("orc_level_pixels", "orc_level_blurred"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]; getattr(L, name).restype = C.c_void_p
        for name in ("orc_level_candidates", "orc_level_keypoints"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
            getattr(L, name).restype = C.c_int
        L.orc_distribute_octree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_void_p, C.c_int]
        L.orc_distribute_octree.restype = C.c_int
        L.orc_descriptor_distance.argtypes = [_u8p, _u8p]; L.orc_descriptor_distance.restype = C.c_int
        L.orc_search_for_initialization.argtypes = (
            [C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p] * 2 +
            [_f32p, C.c_float, C.c_int, C.c_int, _f32p, _i32p])
        L.orc_search_for_initialization.restype = C.c_int
        L.orc_search_by_projection.argtypes = [
            C.c_int, _u8p, _u8p, _f32p, _f32p, _f32p, _i32p, _f32p, _u8p, _i32p,
            C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p, _i32p, _i32p,
            C.c_int, _f32p, _f32p, C.c_float, C.c_float]
        L.orc_search_by_projection.restype = C.c_int
        _lib = L
    return _lib


# ---- primitives -------------------------------------------------------------------------
def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_ptr(src, _u8p), src.shape[1], src.shape[0], src.strides[0],
                               _ptr(dst, _u8p), dw, dh, dw)
    return dst


def copy_make_border(src, border):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h + 2 * border, w + 2 * border), np.uint8)
    lib().orc_copy_make_border_reflect101(_ptr(src, _u8p), w, h, src.strides[0], _ptr(dst, _u8p),
                                          dst.strides[0], border)
    return dst


def fast(img, threshold, nms=True):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = (w // 2 + 2) * (h // 2 + 2) if nms else w * h
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast9_16(_ptr(img, _u8p), w, h, img.strides[0], threshold, int(nms), out.ctypes.data, cap)
    assert n <= cap
    return out[:n]


def gaussian_blur7(img, variant=0):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    dst = np.empty_like(img)
    lib().orc_gaussian_blur7(_ptr(img, _u8p), w, h, img.strides[0], _ptr(dst, _u8p), dst.strides[0], variant)
    return dst


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def sincosf(a):
    s = C.c_float(); c = C.c_float()
    lib().orc_sincosf(float(a), C.byref(s), C.byref(c))
    return s.value, c.value


def distribute_octree(cands, minX, maxX, minY, maxY, N, tie_break=0):
    cands = np.ascontiguousarray(cands, KP_DTYPE)
    cap = N + 4 * 64 + 16
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_distribute_octree(cands.ctypes.data, len(cands), minX, maxX, minY, maxY, N, tie_break,
                                    out.ctypes.data, cap)
    assert n >= 0, n
    return out[:n]


# ---- extractor ---------------------------------------------------------------------------
class OracleExtractor:
    """Mirror of ORB_SLAM2::ORBextractor (I/ORBextractor.h:45-111) backed by the C oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, blur_variant=0):
        self.L = lib()
        self.h = self.L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        assert self.h
        self.L.orc_extractor_set_blur_variant(self.h, blur_variant)
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        try:
            self.L.orc_extractor_destroy(self.h)
        except Exception:
            pass

    def _farr(self, fn, n):
        return np.ctypeslib.as_array(getattr(self.L, fn)(self.h), (n,)).copy()

    @property
    def scale_factors(self): return self._farr("orc_scale_factors", self.nlevels)
    @property
    def inv_scale_factors(self): return self._farr("orc_inv_scale_factors", self.nlevels)
    @property
    def level_sigma2(self): return self._farr("orc_level_sigma2", self.nlevels)
    @property
    def inv_level_sigma2(self): return self._farr("orc_inv_level_sigma2", self.nlevels)
    @property
    def features_per_level(self): return self._farr("orc_features_per_level", self.nlevels)
    @property
    def umax(self): return self._farr("orc_umax", 16)

    def __call__(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 4 * 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_extract(self.h, _ptr(img, _u8p), w, h, img.strides[0], kps.ctypes.data,
                               _ptr(desc, _u8p), cap)
        if n < 0:
            raise ValueError("oracle extract failed: %d" % n)
        return kps[:n].copy(), desc[:n].copy()

    def level_shape(self, l):
        return self.L.orc_level_height(self.h, l), self.L.orc_level_width(self.h, l)

    def level_pixels(self, l):
        h, w = self.level_shape(l)
        p = self.L.orc_level_pixels(self.h, l)
        return np.ctypeslib.as_array(C.cast(p, _u8p), (h, w)).copy()

    def level_blurred(self, l):
        h, w = self.level_shape(l)
        p = self.L.orc_level_blurred(self.h, l)
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, _u8p), (h, w)).copy()

    def _kparr(self, fn, l):
        p = C.c_void_p()
        n = getattr(self.L, fn)(self.h, l, C.byref(p))
        if n == 0:
            return np.zeros(0, KP_DTYPE)
        buf = (C.c_char * (n * 28)).from_address(p.value)
        return np.frombuffer(buf, KP_DTYPE, n).copy()

    def level_candidates(self, l): return self._kparr("orc_level_candidates", l)
    def level_keypoints(self, l): return self._kparr("orc_level_keypoints", l)


# ---- matcher -----------------------------------------------------------------------------
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_ptr(a, _u8p), _ptr(b, _u8p))


def _f(a): return np.ascontiguousarray(a, np.float32)
def _i(a): return np.ascontiguousarray(a, np.int32)
def _b(a): return np.ascontiguousarray(a, np.uint8)


def _bounds(b):
    return np.ascontiguousarray(b, np.float32).reshape(4)


def search_for_initialization(k1, d1, k2, d2, prev_matched, bounds, nnratio=0.9, check_ori=True, window=100):
    """k1/k2: KP_DTYPE arrays (mvKeysUn); prev_matched: (n1,2) float32, updated copy returned;
    bounds = (mnMinX, mnMinY, mnMaxX, mnMaxY)."""
    bnd = _bounds(bounds)
    n1, n2 = len(k1), len(k2)
    k1x, k1y, k1o, k1a = _f(k1["x"]), _f(k1["y"]), _i(k1["octave"]), _f(k1["angle"])
    k2x, k2y, k2o, k2a = _f(k2["x"]), _f(k2["y"]), _i(k2["octave"]), _f(k2["angle"])
    d1 = _b(d1); d2 = _b(d2)
    pm = _f(prev_matched).copy()
    m12 = np.full(max(n1, 1), -1, np.int32)
    n = lib().orc_search_for_initialization(
        n1, _ptr(k1x, _f32p), _ptr(k1y, _f32p), _ptr(k1o, _i32p), _ptr(k1a, _f32p), _ptr(d1, _u8p),
        n2, _ptr(k2x, _f32p), _ptr(k2y, _f32p), _ptr(k2o, _i32p), _ptr(k2a, _f32p), _ptr(d2, _u8p),
        _ptr(bnd, _f32p), nnratio, int(check_ori), window, _ptr(pm, _f32p), _ptr(m12, _i32p))
    return n, m12[:n1], pm


def search_by_projection(mp, kp, kdesc, scale_factors, bounds, nnratio=0.8, th=1.0,
                         kp_mp=None, kp_mp_obs=None):
    """mp: dict of arrays (in_view,bad,x,y,xr,level,viewcos,desc,obs); kp: KP_DTYPE (mvKeysUn)."""
    nmp, n = len(mp["x"]), len(kp)
    kx, ky, ko = _f(kp["x"]), _f(kp["y"]), _i(kp["octave"])
    kur = _f(mp.get("kuright", np.full(n, -1.0, np.float32)))
    kdesc = _b(kdesc)
    kp_mp = np.full(max(n, 1), -1, np.int32) if kp_mp is None else _i(kp_mp).copy()
    kp_mp_obs = np.zeros(max(n, 1), np.int32) if kp_mp_obs is None else _i(kp_mp_obs)
    a = {k: v for k, v in mp.items()}
    iv, bad = _b(a["in_view"]), _b(a["bad"])
    x, y, xr, lv, vc = _f(a["x"]), _f(a["y"]), _f(a["xr"]), _i(a["level"]), _f(a["viewcos"])
    de, ob = _b(a["desc"]), _i(a["obs"])
    sf = _f(scale_factors)
    bnd = _bounds(bounds)
    cnt = lib().orc_search_by_projection(
        nmp, _ptr(iv, _u8p), _ptr(bad, _u8p), _ptr(x, _f32p), _ptr(y, _f32p), _ptr(xr, _f32p),
        _ptr(lv, _i32p), _ptr(vc, _f32p), _ptr(de, _u8p), _ptr(ob, _i32p),
        n, _ptr(kx, _f32p), _ptr(ky, _f32p), _ptr(ko, _i32p), _ptr(kur, _f32p), _ptr(kdesc, _u8p),
        _ptr(kp_mp, _i32p), _ptr(kp_mp_obs, _i32p),
        len(sf), _ptr(sf, _f32p), _ptr(bnd, _f32p), nnratio, th)
    return cnt, kp_mp[:n]


# ---- frame glue ---------------------------------------------------------------------------
def undistort_points(xy, K, dist):
    L = lib()
    L.orc_undistort_points.argtypes = [C.c_int, _f32p, _f32p, _f32p, _f32p]
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    out = np.zeros_like(xy)
    L.orc_undistort_points(len(xy), _ptr(xy, _f32p), _ptr(out, _f32p), _ptr(K, _f32p), _ptr(dist, _f32p))
    return out


def image_bounds(cols, rows, K, dist):
    L = lib()
    L.orc_image_bounds.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p]
    K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.float32)
    L.orc_image_bounds(cols, rows, _ptr(K, _f32p), _ptr(dist, _f32p), _ptr(b, _f32p))
    return b


def _lastframe_args(w, kp_mp, kp_mp_obs, kuright):
    cur = w["cur"]
    n = len(cur)
    a = dict(has_mp=_b(w["has_mp"]), outlier=_b(w["outlier"]), wpos=_f(w["wpos"]), mp_desc=_b(w["mp_desc"]),
             mp_obs=_i(w["mp_obs"]), loct=_i(w["last_octave"]), lang=_f(w["last_angle"]), R=_f(w["Rcw"]), t=_f(w["tcw"]),
             K=_f(w["K"]), kx=_f(cur["x"]), ky=_f(cur["y"]), ko=_i(cur["octave"]), ka=_f(cur["angle"]),
             kur=_f(np.full(n, -1.0, np.float32) if kuright is None else kuright), kd=_b(w["cdesc"]),
             kpmp=np.full(max(n, 1), -1, np.int32) if kp_mp is None else _i(kp_mp).copy(),
             kpobs=np.zeros(max(n, 1), np.int32) if kp_mp_obs is None else _i(kp_mp_obs))
    return a, n


def search_by_projection_last_frame(w, scale_factors, bounds, th=15.0, mode=0, check_ori=True, mbf=40.0,
                                    kp_mp=None, kp_mp_obs=None, kuright=None, fn=None):
    """w: workloads.motion_frame() dict.  mode 0 = octave window, 1 = forward, 2 = backward."""
    L = lib()
    L.orc_search_by_projection_last_frame.argtypes = [
        C.c_int, _u8p, _u8p, _f32p, _u8p, _i32p, _i32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
        C.c_int, _f32p, _f32p, _i32p, _f32p, _f32p, _u8p, _i32p, _i32p,
        C.c_int, _f32p, _f32p, C.c_float, C.c_int, C.c_int]
    L.orc_search_by_projection_last_frame.restype = C.c_int
    a, n = _lastframe_args(w, kp_mp, kp_mp_obs, kuright)
    sf, bnd = _f(scale_factors), _bounds(bounds)
    f = fn or L.orc_search_by_projection_last_frame
    cnt = f(len(a["has_mp"]), _ptr(a["has_mp"], _u8p), _ptr(a["outlier"], _u8p), _ptr(a["wpos"], _f32p),
            _ptr(a["mp_desc"], _u8p), _ptr(a["mp_obs"], _i32p), _ptr(a["loct"], _i32p), _ptr(a["lang"], _f32p),
            _ptr(a["R"], _f32p), _ptr(a["t"], _f32p), _ptr(a["K"], _f32p), mbf,
            n, _ptr(a["kx"], _f32p), _ptr(a["ky"], _f32p), _ptr(a["ko"], _i32p), _ptr(a["ka"], _f32p), _ptr(a["kur"], _f32p),
            _ptr(a["kd"], _u8p), _ptr(a["kpmp"], _i32p), _ptr(a["kpobs"], _i32p),
            len(sf), _ptr(sf, _f32p), _ptr(bnd, _f32p), th, mode, int(check_ori))
    return cnt, a["kpmp"][:n]


def _keyframe_args(w, kp_mp):
    cur = w["cur"]
    n = len(cur)
    a = dict(valid=_b(w["valid"]), wpos=_f(w["wpos"]), mp_desc=_b(w["mp_desc"]), mfmax=_f(w["mf_max"]), mfmin=_f(w["mf_min"]),
             kfang=_f(w["kf_angle"]), R=_f(w["Rcw"]), t=_f(w["tcw"]), Ow=_f(w["Ow"]), K=_f(w["K"]),
             kx=_f(cur["x"]), ky=_f(cur["y"]), ko=_i(cur["octave"]), ka=_f(cur["angle"]), kd=_b(w["cdesc"]),
             kpmp=_i(w["kp_mp"] if kp_mp is None else kp_mp).copy())
    if n == 0:
        a["kpmp"] = np.full(1, -1, np.int32)
    return a, n


def search_by_projection_keyframe(w, scale_factors, bounds, th=10.0, orb_dist=100, check_ori=True, kp_mp=None):
    """w: workloads.relocalisation_frame() dict.  The oracle treats valid == 1 as a usable map point."""
    L = lib()
    L.orc_search_by_projection_keyframe.argtypes = [
        C.c_int, _u8p, _f32p, _u8p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
        C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p, _i32p, C.c_int, _f32p, C.c_float, _f32p,
        C.c_float, C.c_int, C.c_int]
    L.orc_search_by_projection_keyframe.restype = C.c_int
    a, n = _keyframe_args(w, kp_mp)
    ok = _b((a["valid"] == 1).astype(np.uint8))
    sf, bnd = _f(scale_factors), _bounds(bounds)
    cnt = L.orc_search_by_projection_keyframe(
        len(ok), _ptr(ok, _u8p), _ptr(a["wpos"], _f32p), _ptr(a["mp_desc"], _u8p), _ptr(a["mfmax"], _f32p),
        _ptr(a["mfmin"], _f32p), _ptr(a["kfang"], _f32p), _ptr(a["R"], _f32p), _ptr(a["t"], _f32p), _ptr(a["Ow"], _f32p),
        _ptr(a["K"], _f32p), n, _ptr(a["kx"], _f32p), _ptr(a["ky"], _f32p), _ptr(a["ko"], _i32p), _ptr(a["ka"], _f32p),
        _ptr(a["kd"], _u8p), _ptr(a["kpmp"], _i32p), len(sf), _ptr(sf, _f32p), float(w["log_scale"]), _ptr(bnd, _f32p),
        th, orb_dist, int(check_ori))
    return cnt, a["kpmp"][:n]


def logf(x):
    L = lib()
    L.orc_logf.argtypes = [C.c_float]; L.orc_logf.restype = C.c_float
    return L.orc_logf(float(x))


def logf_array(x):
    """orc_logf over a float32 array (loop in C via ctypes is slow; fine for ~1e5 samples)."""
    return np.array([logf(v) for v in np.asarray(x, np.float32)], np.float32)


_u32p = C.POINTER(C.c_uint32)


def _u(a): return np.ascontiguousarray(a, np.uint32)


def _bow_args(w):
    return dict(kv=_b(w["kf_valid"]), kd=_b(w["kf_desc"]), ka=_f(w["kf_angle"]), kn=_u(w["kf_node"]), ks=_i(w["kf_start"]), kf=_u(w["kf_feat"]),
                fd=_b(w["f_desc"]), fa=_f(w["f_angle"]), fn=_u(w["f_node"]), fs=_i(w["f_start"]), ff=_u(w["f_feat"]))


def _bow_call(f, w, nnratio, check_ori, valid):
    a = _bow_args(w)
    nf = len(a["fa"])
    m = np.full(max(nf, 1), -1, np.int32)
    f.argtypes = [C.c_int, _u8p, _u8p, _f32p, C.c_int, _u32p, _i32p, _u32p, C.c_int, _u8p, _f32p, C.c_int, _u32p, _i32p, _u32p,
                  C.c_float, C.c_int, _i32p]
    f.restype = C.c_int
    cnt = f(len(a["kv"]), _ptr(valid(a["kv"]), _u8p), _ptr(a["kd"], _u8p), _ptr(a["ka"], _f32p), len(a["kn"]), _ptr(a["kn"], _u32p),
            _ptr(a["ks"], _i32p), _ptr(a["kf"], _u32p), nf, _ptr(a["fd"], _u8p), _ptr(a["fa"], _f32p), len(a["fn"]),
            _ptr(a["fn"], _u32p), _ptr(a["fs"], _i32p), _ptr(a["ff"], _u32p), nnratio, int(check_ori), _ptr(m, _i32p))
    return cnt, m[:nf]


def search_by_bow(w, nnratio=0.7, check_ori=True):
    """w: workloads.bow_pair() dict.  Returns (nmatches, matches[nf] = key-frame slot or -1)."""
    return _bow_call(lib().orc_search_by_bow, w, nnratio, check_ori, lambda v: _b((v == 1).astype(np.uint8)))


def _bow_kk_call(f, w, nnratio, check_ori, valid):
    a = _bow_args(w)
    v2 = _b(w["f_valid"])
    n1 = len(a["ka"])
    m = np.full(max(n1, 1), -1, np.int32)
    f.argtypes = [C.c_int, _u8p, _u8p, _f32p, C.c_int, _u32p, _i32p, _u32p, C.c_int, _u8p, _u8p, _f32p, C.c_int, _u32p, _i32p, _u32p,
                  C.c_float, C.c_int, _i32p]
    f.restype = C.c_int
    cnt = f(n1, _ptr(valid(a["kv"]), _u8p), _ptr(a["kd"], _u8p), _ptr(a["ka"], _f32p), len(a["kn"]), _ptr(a["kn"], _u32p),
            _ptr(a["ks"], _i32p), _ptr(a["kf"], _u32p), len(a["fa"]), _ptr(valid(v2), _u8p), _ptr(a["fd"], _u8p), _ptr(a["fa"], _f32p),
            len(a["fn"]), _ptr(a["fn"], _u32p), _ptr(a["fs"], _i32p), _ptr(a["ff"], _u32p), nnratio, int(check_ori), _ptr(m, _i32p))
    return cnt, m[:n1]


def search_by_bow_keyframes(w, nnratio=0.75, check_ori=True):
    """w: workloads.bow_pair(..., second_is_keyframe=True) dict.  Returns (nmatches, matches12[n1] = slot of key frame 2 or -1)."""
    return _bow_kk_call(lib().orc_search_by_bow_keyframes, w, nnratio, check_ori, lambda v: _b((v == 1).astype(np.uint8)))


def _tri_sides(w):
    a, b = w["k1"], w["k2"]
    A = dict(mp=_b(a["has_mp"]), d=_b(a["desc"]), x=_f(a["x"]), y=_f(a["y"]), an=_f(a["angle"]), ur=_f(a["u_right"]),
             nd=_u(a["node"]), st=_i(a["start"]), ft=_u(a["feat"]))
    B = dict(mp=_b(b["has_mp"]), d=_b(b["desc"]), x=_f(b["x"]), y=_f(b["y"]), oc=_i(b["octave"]), an=_f(b["angle"]), ur=_f(b["u_right"]),
             nd=_u(b["node"]), st=_i(b["start"]), ft=_u(b["feat"]))
    return A, B


def search_for_triangulation(w, only_stereo=False, check_ori=True):
    """w: workloads.triangulation_pair() dict.  Returns (nmatches, matches12[n1] = index in key frame 2 or -1)."""
    L = lib()
    f = L.orc_search_for_triangulation
    f.argtypes = [C.c_int, _u8p, _u8p, _f32p, _f32p, _f32p, _f32p, C.c_int, _u32p, _i32p, _u32p,
                  C.c_int, _u8p, _u8p, _f32p, _f32p, _i32p, _f32p, _f32p, C.c_int, _u32p, _i32p, _u32p,
                  _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_int, _i32p]
    f.restype = C.c_int
    A, B = _tri_sides(w)
    n1 = len(A["x"])
    m = np.full(max(n1, 1), -1, np.int32)
    F, ep, sf, ls = _f(w["F12"]), _f(w["epipole"]), _f(w["scale_factors"]), _f(w["level_sigma2"])
    cnt = f(n1, _ptr(A["mp"], _u8p), _ptr(A["d"], _u8p), _ptr(A["x"], _f32p), _ptr(A["y"], _f32p), _ptr(A["an"], _f32p), _ptr(A["ur"], _f32p),
            len(A["nd"]), _ptr(A["nd"], _u32p), _ptr(A["st"], _i32p), _ptr(A["ft"], _u32p),
            len(B["x"]), _ptr(B["mp"], _u8p), _ptr(B["d"], _u8p), _ptr(B["x"], _f32p), _ptr(B["y"], _f32p), _ptr(B["oc"], _i32p),
            _ptr(B["an"], _f32p), _ptr(B["ur"], _f32p), len(B["nd"]), _ptr(B["nd"], _u32p), _ptr(B["st"], _i32p), _ptr(B["ft"], _u32p),
            _ptr(F, _f32p), _ptr(ep, _f32p), _ptr(sf, _f32p), _ptr(ls, _f32p), int(only_stereo), int(check_ori), _ptr(m, _i32p))
    return cnt, m[:n1]


def distinctive_descriptor(desc):
    """desc: (n, 32) uint8.  Returns (best index or -1, its median distance)."""
    L = lib()
    L.orc_distinctive_descriptor.argtypes = [C.c_int, _u8p, C.POINTER(C.c_int)]
    L.orc_distinctive_descriptor.restype = C.c_int
    d = _b(desc).reshape(-1, 32)
    med = C.c_int(0)
    idx = L.orc_distinctive_descriptor(len(d), _ptr(d, _u8p), C.byref(med))
    return idx, med.value


def _fuse_call(f, w, bounds, th, valid, with_dist, mode=0, R2=None, t2=None):
    kp = w["kp"]
    n, nmp = len(kp), len(w["valid"])
    a = dict(v=valid(_b(w["valid"])), wp=_f(w["wpos"]), nr=_f(w["normal"]), md=_b(w["mp_desc"]), mx=_f(w["mf_max"]), mn=_f(w["mf_min"]),
             R=_f(w["Rcw"]), t=_f(w["tcw"]), Ow=_f(w["Ow"]), K=_f(w["K"]), kx=_f(kp["x"]), ky=_f(kp["y"]), ko=_i(kp["octave"]),
             ur=_f(w["u_right"]), kd=_b(w["kdesc"]), sf=_f(w["scale_factors"]), il=_f(w["inv_level_sigma2"]), b=_bounds(bounds))
    best = np.full(max(nmp, 1), -1, np.int32); dist = np.full(max(nmp, 1), 256, np.int32)
    args = [nmp, _ptr(a["v"], _u8p), _ptr(a["wp"], _f32p), _ptr(a["nr"], _f32p), _ptr(a["md"], _u8p), _ptr(a["mx"], _f32p), _ptr(a["mn"], _f32p),
            _ptr(a["R"], _f32p), _ptr(a["t"], _f32p), _ptr(a["Ow"], _f32p), _ptr(a["K"], _f32p), float(w["bf"]),
            n, _ptr(a["kx"], _f32p), _ptr(a["ky"], _f32p), _ptr(a["ko"], _i32p), _ptr(a["ur"], _f32p), _ptr(a["kd"], _u8p),
            len(a["sf"]), _ptr(a["sf"], _f32p), _ptr(a["il"], _f32p), float(w["log_scale"]), _ptr(a["b"], _f32p), float(th)]
    f.argtypes = [C.c_int, _u8p, _f32p, _f32p, _u8p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float,
                  C.c_int, _f32p, _f32p, _i32p, _f32p, _u8p, C.c_int, _f32p, _f32p, C.c_float, _f32p, C.c_float]
    f.restype = None
    if with_dist:                     # the oracle's signature: mode, second transform, both outputs
        r2, tt2 = _f(np.zeros(9) if R2 is None else R2), _f(np.zeros(3) if t2 is None else t2)
        f.argtypes = f.argtypes + [C.c_int, _f32p, _f32p, _i32p, _i32p]
        args += [int(mode), _ptr(r2, _f32p), _ptr(tt2, _f32p), _ptr(best, _i32p), _ptr(dist, _i32p)]
    else:                             # the reference harness: Fuse(pKF, vpMapPoints, th) only
        f.argtypes = f.argtypes + [_i32p]
        args.append(_ptr(best, _i32p))
    f(*args)
    return best[:nmp], dist[:nmp]


def fuse_search(w, bounds, th=3.0, mode=0, R2=None, t2=None):
    """w: workloads.fuse_frame() dict.  mode 0 = Fuse(pKF, vpMapPoints, th), 1 = Fuse(pKF, Scw, ...), 2 = a SearchBySim3 leg
    (second transform R2, t2).  Returns (best keypoint index per map point or -1, best distance)."""
    return _fuse_call(lib().orc_fuse_search, w, bounds, th, lambda v: _b((v == 1).astype(np.uint8)), True, mode, R2, t2)


def search_by_sim3(w, bounds, th=7.5):
    """w: workloads.sim3_pair() dict.  SearchBySim3 from its two legs + the agreement test.
    Returns (nFound, matches12[n1] = slot of key frame 2 or -1)."""
    L = lib()
    m1, _ = fuse_search(w["leg1"], bounds, th, 2, w["sR21"], w["t21"])
    m2, _ = fuse_search(w["leg2"], bounds, th, 2, w["R12"], w["t12"])
    L.orc_sim3_agreement.argtypes = [C.c_int, _i32p, C.c_int, _i32p, _i32p]
    L.orc_sim3_agreement.restype = C.c_int
    a, b = _i(m1) if len(m1) else np.zeros(1, np.int32), _i(m2) if len(m2) else np.zeros(1, np.int32)
    out = np.full(max(len(m1), 1), -1, np.int32)
    found = L.orc_sim3_agreement(len(m1), _ptr(a, _i32p), len(m2), _ptr(b, _i32p), _ptr(out, _i32p))
    return found, out[:len(m1)]


def _proj_sim3_call(f, w, bounds, th, valid, matched):
    kp = w["kp"]
    n, nmp = len(kp), len(w["valid"])
    a = dict(v=valid(_b(w["valid"])), wp=_f(w["wpos"]), nr=_f(w["normal"]), md=_b(w["mp_desc"]), mx=_f(w["mf_max"]), mn=_f(w["mf_min"]),
             R=_f(w["Rcw"]), t=_f(w["tcw"]), Ow=_f(w["Ow"]), K=_f(w["K"]), kx=_f(kp["x"]), ky=_f(kp["y"]), ko=_i(kp["octave"]),
             kd=_b(w["kdesc"]), sf=_f(w["scale_factors"]), b=_bounds(bounds))
    m = np.full(max(n, 1), -1, np.int32)
    if matched is not None and n:
        m[:n] = matched
    f.argtypes = [C.c_int, _u8p, _f32p, _f32p, _u8p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
                  C.c_int, _f32p, _f32p, _i32p, _u8p, C.c_int, _f32p, C.c_float, _f32p, C.c_int, _i32p]
    f.restype = C.c_int
    cnt = f(nmp, _ptr(a["v"], _u8p), _ptr(a["wp"], _f32p), _ptr(a["nr"], _f32p), _ptr(a["md"], _u8p), _ptr(a["mx"], _f32p), _ptr(a["mn"], _f32p),
            _ptr(a["R"], _f32p), _ptr(a["t"], _f32p), _ptr(a["Ow"], _f32p), _ptr(a["K"], _f32p), n, _ptr(a["kx"], _f32p), _ptr(a["ky"], _f32p),
            _ptr(a["ko"], _i32p), _ptr(a["kd"], _u8p), len(a["sf"]), _ptr(a["sf"], _f32p), float(w["log_scale"]), _ptr(a["b"], _f32p), int(th),
            _ptr(m, _i32p))
    return cnt, m[:n]


def search_by_projection_sim3(w, bounds, th=10, matched=None):
    """w: workloads.fuse_frame() dict (valid == 1 = usable candidate).  Returns (nmatches, matched[n])."""
    return _proj_sim3_call(lib().orc_search_by_projection_sim3, w, bounds, th, lambda v: _b((v == 1).astype(np.uint8)), matched)


_f64p = C.POINTER(C.c_double)


def _bow_out(n):
    m = max(n, 1)
    return dict(bn=C.c_int32(0), bw=np.zeros(m, np.uint32), bv=np.zeros(m, np.float64), fn=C.c_int32(0), fnode=np.zeros(m, np.uint32),
                fstart=np.zeros(m + 1, np.int32), ffeat=np.zeros(m, np.uint32))


def _bow_result(o):
    nb, nf = o["bn"].value, o["fn"].value
    return dict(word=o["bw"][:nb].copy(), value=o["bv"][:nb].copy(), node=o["fnode"][:nf].copy(), start=o["fstart"][:nf + 1].copy(),
                feat=o["ffeat"][:o["fstart"][nf]].copy())


def bow_transform(voc, desc, levelsup=4):
    """voc: workloads.synthetic_vocabulary() dict; desc: (n, 32) uint8.  Returns dict(word, value, node, start, feat)."""
    L = lib()
    f = L.orc_bow_transform
    f.argtypes = [C.c_int, C.c_int, _i32p, _i32p, _u8p, _i32p, _f64p, C.c_int, _u8p, C.c_int,
                  C.POINTER(C.c_int32), _u32p, _f64p, C.POINTER(C.c_int32), _u32p, _i32p, _u32p]
    f.restype = None
    d = _b(desc).reshape(-1, 32)
    cs, ch, nd, wi = _i(voc["child_start"]), _i(voc["children"]), _b(voc["desc"]), _i(voc["word_id"])
    wt = np.ascontiguousarray(voc["weight"], np.float64)
    o = _bow_out(len(d))
    f(len(voc["parent"]), voc["L"], _ptr(cs, _i32p), _ptr(ch, _i32p), _ptr(nd, _u8p), _ptr(wi, _i32p), _ptr(wt, _f64p), len(d), _ptr(d, _u8p),
      int(levelsup), C.byref(o["bn"]), _ptr(o["bw"], _u32p), _ptr(o["bv"], _f64p), C.byref(o["fn"]), _ptr(o["fnode"], _u32p),
      _ptr(o["fstart"], _i32p), _ptr(o["ffeat"], _u32p))
    return _bow_result(o)


def _pyr_args(pyr):
    """list of 2-D uint8 level images -> (keep-alive list, pointer array, pitch array)."""
    lv = [np.ascontiguousarray(p, np.uint8) for p in pyr]
    ptrs = (_u8p * len(lv))(*[_ptr(p, _u8p) for p in lv])
    pitch = np.array([p.strides[0] for p in lv], np.int32)
    return lv, ptrs, pitch


def compute_stereo_matches(kl, dl, kr, dr, scale, inv_scale, lpyr, rpyr, mb, mbf):
    """Frame::ComputeStereoMatches.  kl / kr: keypoint records (x, y, octave); dl / dr: (n, 32) descriptors;
    lpyr / rpyr: pyramid levels without border.  Returns (u_right, depth, kept, skipped)."""
    f = lib().orc_compute_stereo_matches
    f.argtypes = [C.c_int, _f32p, _f32p, _i32p, _u8p, C.c_int, _f32p, _f32p, _i32p, _u8p, C.c_int, _f32p, _f32p,
                  C.POINTER(_u8p), _i32p, C.POINTER(_u8p), _i32p, _i32p, _i32p, C.c_float, C.c_float, _f32p, _f32p, _i32p]
    f.restype = C.c_int
    n, nr = len(kl), len(kr)
    a = [_f(kl["x"]), _f(kl["y"]), _i(kl["octave"]), _b(dl) if n else np.zeros((1, 32), np.uint8)]
    b = [_f(kr["x"]), _f(kr["y"]), _i(kr["octave"]), _b(dr) if nr else np.zeros((1, 32), np.uint8)]
    sc, isc = _f(scale), _f(inv_scale)
    lk, lp, lpitch = _pyr_args(lpyr)
    rk, rp, rpitch = _pyr_args(rpyr)
    lw = np.array([p.shape[1] for p in lk], np.int32); lh = np.array([p.shape[0] for p in lk], np.int32)
    ur = np.zeros(max(n, 1), np.float32); dep = np.zeros(max(n, 1), np.float32); skipped = np.zeros(1, np.int32)
    kept = f(n, _ptr(a[0], _f32p), _ptr(a[1], _f32p), _ptr(a[2], _i32p), _ptr(a[3], _u8p),
             nr, _ptr(b[0], _f32p), _ptr(b[1], _f32p), _ptr(b[2], _i32p), _ptr(b[3], _u8p), len(sc), _ptr(sc, _f32p), _ptr(isc, _f32p),
             lp, _ptr(lpitch, _i32p), rp, _ptr(rpitch, _i32p), _ptr(lw, _i32p), _ptr(lh, _i32p), mb, mbf,
             _ptr(ur, _f32p), _ptr(dep, _f32p), _ptr(skipped, _i32p))
    return ur[:n], dep[:n], kept, int(skipped[0])
