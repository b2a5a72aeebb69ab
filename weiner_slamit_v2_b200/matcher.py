"""Host-side mirror of ORB_SLAM2::ORBmatcher (I/ORBmatcher.h:37-102) over the C ABI, for the three
in-scope entry points: DescriptorDistance, SearchForInitialization, SearchByProjection(Frame&,
vector<MapPoint*>&, th).  Frame / MapPoint pointer graphs are flattened to structure-of-arrays
views exactly as a C++ maintainer's shim would; results are written back into the
same fields the reference mutates (vnMatches12, vbPrevMatched, Frame::mvpMapPoints)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import FrameView, KP_DTYPE, MapPointView, check

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30       # S/ORBmatcher.cc:37-39


class Frame:
    """The fields of ORB_SLAM2::Frame the matcher reads (I/Frame.h): mvKeysUn, mDescriptors, mvuRight,
    mvpMapPoints (as indices into the map-point list, -1 = none), mvScaleFactors, image bounds."""

    def __init__(self, keys_un, descriptors, width, height, scale_factors=None, u_right=None, bounds=None):
        self.mvKeysUn = np.ascontiguousarray(keys_un, KP_DTYPE)
        self.mDescriptors = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        assert len(self.mvKeysUn) == len(self.mDescriptors)
        self.N = len(self.mvKeysUn)
        self.width, self.height = int(width), int(height)
        # mnMinX, mnMinY, mnMaxX, mnMaxY (S/Frame.cc:561-589): the image rectangle unless lens distortion moves it
        self.bounds = np.array([0, 0, width, height] if bounds is None else bounds, np.float32)
        self.mvScaleFactors = None if scale_factors is None else np.ascontiguousarray(scale_factors, np.float32)
        self.mvuRight = np.full(self.N, -1.0, np.float32) if u_right is None else np.ascontiguousarray(u_right, np.float32)
        self.mvpMapPoints = np.full(self.N, -1, np.int32)
        self.mvpMapPointObs = np.zeros(self.N, np.int32)     # Observations() of foreign map points (-2 entries)


class MapPoints:
    """vector<MapPoint*> flattened: the 'variables used by the tracking' (I/MapPoint.h:96-104) plus
    GetDescriptor(), isBad(), Observations()."""

    def __init__(self, proj_x, proj_y, level, view_cos, desc, in_view=None, bad=None, proj_xr=None, obs=None):
        n = len(proj_x)
        self.n = n
        self.mTrackProjX = np.ascontiguousarray(proj_x, np.float32)
        self.mTrackProjY = np.ascontiguousarray(proj_y, np.float32)
        self.mTrackProjXR = np.zeros(n, np.float32) if proj_xr is None else np.ascontiguousarray(proj_xr, np.float32)
        self.mnTrackScaleLevel = np.ascontiguousarray(level, np.int32)
        self.mTrackViewCos = np.ascontiguousarray(view_cos, np.float32)
        self.descriptor = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.mbTrackInView = np.ones(n, np.uint8) if in_view is None else np.ascontiguousarray(in_view, np.uint8)
        self.bad = np.zeros(n, np.uint8) if bad is None else np.ascontiguousarray(bad, np.uint8)
        self.observations = np.ones(n, np.int32) if obs is None else np.ascontiguousarray(obs, np.int32)


def _pack(arrs, stride, dtype, tail=()):
    out = np.zeros((len(arrs), stride) + tuple(tail), dtype)
    for i, a in enumerate(arrs):
        out[i, :len(a)] = a
    return out


def _device_copy(view, arrays, device, keep):
    """A copy of the ctypes view `view` whose pointer fields address device copies of the host arrays they pointed to
    (`arrays`: the numpy arrays the view was built from).  torch is plumbing: device memory only.  Used by bench.py to time
    an entry point with on_device = 1 on exactly the data its host-buffer form was checked on."""
    import torch
    by_addr = {a.ctypes.data: a for a in arrays if a is not None}
    out = type(view)()
    for name, ctype in view._fields_:
        v = getattr(view, name)
        if ctype is _lib.vp and v in by_addr:
            t = torch.from_numpy(by_addr[v]).to(torch.device("cuda", device))
            keep.append(t)
            setattr(out, name, t.data_ptr())
        else:
            setattr(out, name, v)
    return out


def _device_array(a, device, keep):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).to(torch.device("cuda", device))
    keep.append(t)
    return t.data_ptr()


def _device_time(stream, device, fn, reps):
    """ms per call of fn (which queues work on `stream`), CUDA events on that stream."""
    import torch
    st = torch.cuda.ExternalStream(stream, device=device)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _bow_side(sides, valid=None):
    """Pack one side of a SearchByBoW-style call.  sides: list of dicts with desc (n,32), angle (n), node, start, feat
    (a flattened DBoW2::FeatureVector); valid: optional list of (n,) uint8.  Returns (arrays kept alive, BowView, stride)."""
    from ._lib import BowView
    items = len(sides)
    n = np.array([len(p["angle"]) for p in sides], np.int32)
    nn = np.array([len(p["node"]) for p in sides], np.int32)
    s, ns = max(1, int(n.max())), max(1, int(nn.max()))
    start = np.zeros((items, ns + 1), np.int32)
    for i, p in enumerate(sides):
        st = np.asarray(p["start"], np.int32)
        start[i, :len(st)] = st
        start[i, len(st):] = st[-1] if len(st) else 0
    a = dict(n=n, nn=nn, desc=_pack([p["desc"] for p in sides], s, np.uint8, (32,)), ang=_pack([p["angle"] for p in sides], s, np.float32),
             node=_pack([p["node"] for p in sides], ns, np.uint32), start=start, feat=_pack([p["feat"] for p in sides], s, np.uint32))
    if valid is not None:
        a["valid"] = _pack(valid, s, np.uint8)
    v = BowView(a["n"].ctypes.data, a["desc"].ctypes.data, a["ang"].ctypes.data, a["valid"].ctypes.data if valid is not None else None,
                a["nn"].ctypes.data, a["node"].ctypes.data, a["start"].ctypes.data, a["feat"].ctypes.data, s, ns)
    return a, v, s


def _frame_view(frames, keep):
    """Pack a list of Frame into padded SoA arrays + the ctypes view (arrays kept alive in `keep`)."""
    stride = max(1, max(f.N for f in frames))
    n = np.array([f.N for f in frames], np.int32)
    x = _pack([f.mvKeysUn["x"] for f in frames], stride, np.float32)
    y = _pack([f.mvKeysUn["y"] for f in frames], stride, np.float32)
    o = _pack([f.mvKeysUn["octave"] for f in frames], stride, np.int32)
    a = _pack([f.mvKeysUn["angle"] for f in frames], stride, np.float32)
    d = _pack([f.mDescriptors for f in frames], stride, np.uint8, (32,))
    keep += [n, x, y, o, a, d]
    return FrameView(n.ctypes.data, x.ctypes.data, y.ctypes.data, o.ctypes.data, a.ctypes.data, d.ctypes.data, stride), stride


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = TH_HIGH, TH_LOW, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True, device=0, max_items=1, max_points=4096):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self._L = _lib.load()
        self._h = _lib.vp()
        self.device, self.max_items, self.max_points = device, max_items, max_points
        check(self._L.orbb200_matcher_create(max_items, max_points, device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbb200_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ensure(self, items, points):
        if items > self.max_items or points > self.max_points:
            self.close()
            self.max_items, self.max_points = max(items, self.max_items), max(points, self.max_points)
            self._h = _lib.vp()
            check(self._L.orbb200_matcher_create(self.max_items, self.max_points, self.device, C.byref(self._h)))

    @property
    def stream(self): return self._L.orbb200_matcher_stream(self._h)
    @property
    def last_launches(self): return self._L.orbb200_matcher_last_launches(self._h)
    def sync(self): check(self._L.orbb200_matcher_sync(self._h))

    # ---- DescriptorDistance (static in the reference; S/ORBmatcher.cc:1651-1667) ----
    def DescriptorDistance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.zeros(len(a), np.int32)
        check(self._L.orbb200_descriptor_distance(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return int(out[0]) if len(out) == 1 else out

    # ---- SearchForInitialization (S/ORBmatcher.cc:409-524) ----
    def SearchForInitialization(self, F1, F2, vbPrevMatched, vnMatches12=None, windowSize=10):
        """Returns nmatches; vbPrevMatched ((n1,2) float32) is updated in place; vnMatches12 is returned
        through the list/array argument when given, and always as the second return value."""
        n, m12, pm = self.search_for_initialization_batch([F1], [F2], [vbPrevMatched], windowSize)
        vbPrevMatched[...] = pm[0]
        if vnMatches12 is not None:
            vnMatches12[:] = m12[0].tolist() if isinstance(vnMatches12, list) else m12[0]
        return int(n[0]), m12[0]

    def search_for_initialization_batch(self, F1s, F2s, prev_matched, windowSize=10):
        items = len(F1s)
        keep = []
        v1, s1 = _frame_view(F1s, keep)
        v2, s2 = _frame_view(F2s, keep)
        self._ensure(items, max(s1, s2))
        bnd = F2s[0].bounds
        assert all(np.array_equal(f.bounds, bnd) for f in F2s), "one image geometry per call"
        pm = _pack([np.asarray(p, np.float32).reshape(-1, 2) for p in prev_matched], s1, np.float32, (2,))
        m12 = np.full((items, s1), -1, np.int32)
        nm = np.zeros(items, np.int32)
        check(self._L.orbb200_search_for_initialization(self._h, items, C.byref(v1), C.byref(v2), bnd.ctypes.data, self.mfNNratio,
                                                        int(self.mbCheckOrientation), int(windowSize), pm.ctypes.data,
                                                        m12.ctypes.data, nm.ctypes.data, 0))
        return nm, [m12[i, :F1s[i].N] for i in range(items)], [pm[i, :F1s[i].N] for i in range(items)]

    # ---- SearchByProjection(Frame&, vector<MapPoint*>&, th) (S/ORBmatcher.cc:47-131) ----
    def SearchByProjection(self, F, vpMapPoints, th=3.0):
        n = self.search_by_projection_batch([F], [vpMapPoints], th)
        return int(n[0])

    def search_by_projection_batch(self, frames, mappoints, th=3.0):
        """Assignments are written into each frame's mvpMapPoints (index of the map point, like
        F.mvpMapPoints[bestIdx] = pMP)."""
        items = len(frames)
        keep = []
        fv, s = _frame_view(frames, keep)
        ms = max(1, max(m.n for m in mappoints))
        self._ensure(items, max(s, ms))
        bnd = frames[0].bounds
        sf = frames[0].mvScaleFactors
        assert sf is not None, "Frame.mvScaleFactors is required"
        ur = _pack([f.mvuRight for f in frames], s, np.float32)
        kpmp = _pack([f.mvpMapPoints for f in frames], s, np.int32)
        kpobs = _pack([f.mvpMapPointObs for f in frames], s, np.int32)
        mn = np.array([m.n for m in mappoints], np.int32)
        a = dict(
            iv=_pack([m.mbTrackInView for m in mappoints], ms, np.uint8), bad=_pack([m.bad for m in mappoints], ms, np.uint8),
            x=_pack([m.mTrackProjX for m in mappoints], ms, np.float32), y=_pack([m.mTrackProjY for m in mappoints], ms, np.float32),
            xr=_pack([m.mTrackProjXR for m in mappoints], ms, np.float32), lv=_pack([m.mnTrackScaleLevel for m in mappoints], ms, np.int32),
            vc=_pack([m.mTrackViewCos for m in mappoints], ms, np.float32), de=_pack([m.descriptor for m in mappoints], ms, np.uint8, (32,)),
            ob=_pack([m.observations for m in mappoints], ms, np.int32))
        mv = MapPointView(mn.ctypes.data, a["iv"].ctypes.data, a["bad"].ctypes.data, a["x"].ctypes.data, a["y"].ctypes.data,
                          a["xr"].ctypes.data, a["lv"].ctypes.data, a["vc"].ctypes.data, a["de"].ctypes.data,
                          a["ob"].ctypes.data, ms)
        nm = np.zeros(items, np.int32)
        check(self._L.orbb200_search_by_projection(self._h, items, C.byref(fv), ur.ctypes.data, C.byref(mv), kpmp.ctypes.data,
                                                   kpobs.ctypes.data, sf.ctypes.data, len(sf), bnd.ctypes.data, self.mfNNratio,
                                                   float(th), nm.ctypes.data, 0))
        for i, f in enumerate(frames):
            f.mvpMapPoints[:] = kpmp[i, :f.N]
        return nm


    # ---- Frame glue: Frame::UndistortKeyPoints / ComputeImageBounds on the device ----
    def undistort_points(self, xy, K, dist):
        """cv::undistortPoints(src, dst, K, dist, Mat(), K) for an (n, 2) float32 array."""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32)
        out = np.zeros_like(xy)
        check(self._L.orbb200_undistort_points(self._h, xy.ctypes.data, out.ctypes.data, len(xy), K.ctypes.data, dist.ctypes.data))
        return out

    def image_bounds(self, cols, rows, K, dist):
        K = np.ascontiguousarray(K, np.float32); dist = np.ascontiguousarray(dist, np.float32)
        b = np.zeros(4, np.float32)
        check(self._L.orbb200_image_bounds(self._h, cols, rows, K.ctypes.data, dist.ctypes.data, b.ctypes.data))
        return b

    # ---- SearchByBoW(pKF, F, vpMapPointMatches) (S/ORBmatcher.cc:161-292), scope row N3 ----
    def search_by_bow_batch(self, pairs, keyframes=False):
        """keyframes=True: SearchByBoW(pKF1, pKF2, vpMatches12) (S/ORBmatcher.cc:526-659): the second side carries
        f_valid too and the result is indexed by the first side (slot of key frame 2 or -1).
        pairs: list of dicts (workloads.bow_pair layout): kf_valid (1 = usable map point), kf_desc, kf_angle,
        kf_node / kf_start / kf_feat (flattened pKF->mFeatVec), f_desc, f_angle, f_node / f_start / f_feat (F.mFeatVec).
        Returns (nmatches (items,), [matches per item: key-frame slot per frame keypoint or -1])."""
        items = len(pairs)

        def side(pfx, with_valid):
            sub = [dict(desc=p[pfx + "_desc"], angle=p[pfx + "_angle"], node=p[pfx + "_node"], start=p[pfx + "_start"], feat=p[pfx + "_feat"])
                   for p in pairs]
            return _bow_side(sub, [(np.asarray(p[pfx + "_valid"]) == 1).astype(np.uint8) for p in pairs] if with_valid else None)
        ka, kv, ks = side("kf", True)
        fa, fv, fs = side("f", keyframes)
        self._ensure(items, max(ks, fs))
        m = np.full((items, ks if keyframes else fs), -1, np.int32)
        nm = np.zeros(items, np.int32)
        fn = self._L.orbb200_search_by_bow_keyframes if keyframes else self._L.orbb200_search_by_bow
        check(fn(self._h, items, C.byref(kv), C.byref(fv), float(self.mfNNratio), int(self.mbCheckOrientation),
                 m.ctypes.data, nm.ctypes.data, 0))
        return nm, [m[i, :(ka if keyframes else fa)["n"][i]] for i in range(items)]

    # ---- the search of Fuse(pKF, vpMapPoints, th) (S/ORBmatcher.cc:829-948), scope row N3 ----
    def fuse_search_batch(self, ws, bounds, th=3.0, mode=0, R2=None, t2=None, device_reps=0):
        """mode 0 = Fuse(pKF, vpMapPoints, th), 1 = Fuse(pKF, Scw, ...), 2 = a SearchBySim3 leg (R2, t2: per-item second
        similarity).  ws: list of workloads.fuse_frame()-layout dicts (valid == 1 marks usable candidates; kp / kdesc / u_right = the
        key frame; Rcw, tcw, Ow, K, bf, scale_factors, inv_level_sigma2, log_scale).  Returns per item
        (best keypoint index per candidate or -1, smallest distance seen)."""
        from ._lib import FrameView, FusePointsView
        items = len(ws)
        nk = np.array([len(w["kp"]) for w in ws], np.int32)
        nm = np.array([len(w["valid"]) for w in ws], np.int32)
        s, ms = max(1, int(nk.max())), max(1, int(nm.max()))
        self._ensure(items, max(s, ms))
        k = dict(x=_pack([w["kp"]["x"] for w in ws], s, np.float32), y=_pack([w["kp"]["y"] for w in ws], s, np.float32),
                 o=_pack([w["kp"]["octave"] for w in ws], s, np.int32), d=_pack([w["kdesc"] for w in ws], s, np.uint8, (32,)),
                 ur=_pack([w["u_right"] for w in ws], s, np.float32))
        fv = FrameView(nk.ctypes.data, k["x"].ctypes.data, k["y"].ctypes.data, k["o"].ctypes.data, None, k["d"].ctypes.data, s)
        a = dict(v=_pack([(np.asarray(w["valid"]) == 1).astype(np.uint8) for w in ws], ms, np.uint8),
                 wp=_pack([np.asarray(w["wpos"], np.float32).reshape(-1, 3) for w in ws], ms, np.float32, (3,)),
                 nr=_pack([np.asarray(w["normal"], np.float32).reshape(-1, 3) for w in ws], ms, np.float32, (3,)),
                 md=_pack([w["mp_desc"] for w in ws], ms, np.uint8, (32,)), mx=_pack([w["mf_max"] for w in ws], ms, np.float32),
                 mn=_pack([w["mf_min"] for w in ws], ms, np.float32))
        pv = FusePointsView(nm.ctypes.data, a["v"].ctypes.data, a["wp"].ctypes.data, a["nr"].ctypes.data, a["md"].ctypes.data,
                            a["mx"].ctypes.data, a["mn"].ctypes.data, ms)
        R = np.ascontiguousarray(np.stack([np.asarray(w["Rcw"], np.float32).reshape(9) for w in ws]))
        t = np.ascontiguousarray(np.stack([np.asarray(w["tcw"], np.float32).reshape(3) for w in ws]))
        Ow = np.ascontiguousarray(np.stack([np.asarray(w["Ow"], np.float32).reshape(3) for w in ws]))
        K = np.ascontiguousarray(ws[0]["K"], np.float32)
        sf = np.ascontiguousarray(ws[0]["scale_factors"], np.float32)
        il = np.ascontiguousarray(ws[0]["inv_level_sigma2"], np.float32)
        bnd = np.ascontiguousarray(bounds, np.float32)
        best = np.full((items, ms), -1, np.int32); dist = np.full((items, ms), 256, np.int32)
        r2 = np.ascontiguousarray(np.stack([np.asarray(r, np.float32).reshape(9) for r in R2])) if R2 is not None else None
        tt2 = np.ascontiguousarray(np.stack([np.asarray(v, np.float32).reshape(3) for v in t2])) if t2 is not None else None
        check(self._L.orbb200_fuse_search(self._h, items, C.byref(fv), k["ur"].ctypes.data, C.byref(pv), R.ctypes.data, t.ctypes.data,
                                          Ow.ctypes.data, K.ctypes.data, float(ws[0]["bf"]), sf.ctypes.data, il.ctypes.data, len(sf),
                                          float(ws[0]["log_scale"]), bnd.ctypes.data, float(th), int(mode),
                                          r2.ctypes.data if r2 is not None else None, tt2.ctypes.data if tt2 is not None else None,
                                          best.ctypes.data, dist.ctypes.data, 0))
        if device_reps:          # the same call on device-resident copies of the same data, CUDA events on the matcher's stream
            keep = []
            dfv = _device_copy(fv, [nk] + list(k.values()), self.device, keep)
            dpv = _device_copy(pv, [nm] + list(a.values()), self.device, keep)
            d = lambda x: _device_array(x, self.device, keep) if x is not None else None
            dur, dR, dT, dOw, dr2, dt2, dsf, dil = d(k["ur"]), d(R), d(t), d(Ow), d(r2), d(tt2), d(sf), d(il)
            dbest, ddist = d(best), d(dist)
            self.last_device_ms = _device_time(self._L.orbb200_matcher_stream(self._h), self.device, lambda: check(
                self._L.orbb200_fuse_search(self._h, items, C.byref(dfv), dur, C.byref(dpv), dR, dT, dOw, K.ctypes.data, float(ws[0]["bf"]),
                                            dsf, dil, len(sf), float(ws[0]["log_scale"]), bnd.ctypes.data, float(th),
                                            int(mode), dr2, dt2, dbest, ddist, 1)), device_reps)
            assert np.array_equal(keep[-2].cpu().numpy(), best), "device-resident call differs from the host-buffer call"
        return [(best[i, :nm[i]], dist[i, :nm[i]]) for i in range(items)]

    # ---- SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (S/ORBmatcher.cc:294-407), scope row N3 ----
    def search_by_projection_sim3_batch(self, ws, bounds, th=10, matched=None):
        """ws: list of workloads.fuse_frame()-layout dicts (valid == 1 = usable candidate, pose = the decomposed Scw).
        matched: optional list of per-item int32 arrays (vpMatched as indices, -1 free).  Returns (nmatches, [matched])."""
        from ._lib import FrameView, FusePointsView
        items = len(ws)
        nk = np.array([len(w["kp"]) for w in ws], np.int32)
        nm = np.array([len(w["valid"]) for w in ws], np.int32)
        s, ms = max(1, int(nk.max())), max(1, int(nm.max()))
        self._ensure(items, max(s, ms))
        k = dict(x=_pack([w["kp"]["x"] for w in ws], s, np.float32), y=_pack([w["kp"]["y"] for w in ws], s, np.float32),
                 o=_pack([w["kp"]["octave"] for w in ws], s, np.int32), d=_pack([w["kdesc"] for w in ws], s, np.uint8, (32,)))
        fv = FrameView(nk.ctypes.data, k["x"].ctypes.data, k["y"].ctypes.data, k["o"].ctypes.data, None, k["d"].ctypes.data, s)
        a = dict(v=_pack([(np.asarray(w["valid"]) == 1).astype(np.uint8) for w in ws], ms, np.uint8),
                 wp=_pack([np.asarray(w["wpos"], np.float32).reshape(-1, 3) for w in ws], ms, np.float32, (3,)),
                 nr=_pack([np.asarray(w["normal"], np.float32).reshape(-1, 3) for w in ws], ms, np.float32, (3,)),
                 md=_pack([w["mp_desc"] for w in ws], ms, np.uint8, (32,)), mx=_pack([w["mf_max"] for w in ws], ms, np.float32),
                 mn=_pack([w["mf_min"] for w in ws], ms, np.float32))
        pv = FusePointsView(nm.ctypes.data, a["v"].ctypes.data, a["wp"].ctypes.data, a["nr"].ctypes.data, a["md"].ctypes.data,
                            a["mx"].ctypes.data, a["mn"].ctypes.data, ms)
        R = np.ascontiguousarray(np.stack([np.asarray(w["Rcw"], np.float32).reshape(9) for w in ws]))
        t = np.ascontiguousarray(np.stack([np.asarray(w["tcw"], np.float32).reshape(3) for w in ws]))
        Ow = np.ascontiguousarray(np.stack([np.asarray(w["Ow"], np.float32).reshape(3) for w in ws]))
        K = np.ascontiguousarray(ws[0]["K"], np.float32)
        sf = np.ascontiguousarray(ws[0]["scale_factors"], np.float32)
        bnd = np.ascontiguousarray(bounds, np.float32)
        mt = np.full((items, s), -1, np.int32)
        if matched is not None:
            for i, mm in enumerate(matched):
                mt[i, :len(mm)] = mm
        cnt = np.zeros(items, np.int32)
        check(self._L.orbb200_search_by_projection_sim3(self._h, items, C.byref(fv), C.byref(pv), R.ctypes.data, t.ctypes.data, Ow.ctypes.data,
                                                        K.ctypes.data, sf.ctypes.data, len(sf), float(ws[0]["log_scale"]), bnd.ctypes.data,
                                                        int(th), mt.ctypes.data, cnt.ctypes.data, 0))
        return cnt, [mt[i, :nk[i]] for i in range(items)]

    # ---- SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (S/ORBmatcher.cc:1106-1330), scope row N3 ----
    def search_by_sim3_batch(self, ws, bounds, th=7.5):
        """ws: list of workloads.sim3_pair()-layout dicts (leg1 / leg2 in fuse_frame layout, sR21, t21, R12 (= sR12), t12).
        Two device searches and the agreement test of :1316-1328.  Returns (nFound (items,), [matches12 per item])."""
        leg1 = self.fuse_search_batch([w["leg1"] for w in ws], bounds, th, 2, [w["sR21"] for w in ws], [w["t21"] for w in ws])
        leg2 = self.fuse_search_batch([w["leg2"] for w in ws], bounds, th, 2, [w["R12"] for w in ws], [w["t12"] for w in ws])
        found, out = np.zeros(len(ws), np.int32), []
        for i in range(len(ws)):
            m1, m2 = leg1[i][0], leg2[i][0]
            ok = (m1 >= 0)
            ok[ok] = m2[m1[ok]] == np.nonzero(ok)[0]
            out.append(np.where(ok, m1, -1).astype(np.int32))
            found[i] = int(ok.sum())
        return found, out

    # ---- MapPoint::ComputeDistinctiveDescriptors (S/MapPoint.cc:248-313), scope row N4 ----
    def distinctive_descriptors_batch(self, observed):
        """observed: list of (n_p, 32) uint8 arrays, one per map point (descriptors of its non-bad key frames in
        observation-map order).  Returns (best index per map point or -1, median distance of the winner)."""
        items = len(observed)
        off = np.zeros(items + 1, np.int32)
        off[1:] = np.cumsum([len(o) for o in observed])
        total = int(off[-1])
        desc = np.zeros((max(total, 1), 32), np.uint8)
        if total:
            desc[:total] = np.concatenate([np.asarray(o, np.uint8).reshape(-1, 32) for o in observed])
        best = np.zeros(items, np.int32); med = np.zeros(items, np.int32)
        check(self._L.orbb200_distinctive_descriptors(self._h, items, off.ctypes.data, desc.ctypes.data, total, best.ctypes.data,
                                                      med.ctypes.data, 0))
        return best, med

    # ---- Frame::ComputeStereoMatches (S/Frame.cc:591-763), scope row N4 ----
    def compute_stereo_matches_batch(self, lefts, rights, lpyr, rpyr, scale_factors, inv_scale_factors, mb, mbf):
        """lefts / rights: per stereo pair (keypoint records, (n, 32) descriptors) = mvKeys / mDescriptors and
        mvKeysRight / mDescriptorsRight.  lpyr / rpyr: either a _lib.PyramidView of device memory
        (ORBextractor.pyramid_view()) or, per pair, the list of level images (host arrays, no border).
        Returns (mvuRight, mvDepth) per pair and the number of stereo matches per pair."""
        items = len(lefts)
        keep = []

        def side(fr):
            stride = max(1, max(len(k) for k, _ in fr))
            n = np.array([len(k) for k, _ in fr], np.int32)
            x = _pack([k["x"] for k, _ in fr], stride, np.float32)
            y = _pack([k["y"] for k, _ in fr], stride, np.float32)
            o = _pack([k["octave"] for k, _ in fr], stride, np.int32)
            d = _pack([np.asarray(dd, np.uint8).reshape(-1, 32) for _, dd in fr], stride, np.uint8, (32,))
            keep.extend([n, x, y, o, d])
            return FrameView(n.ctypes.data, x.ctypes.data, y.ctypes.data, o.ctypes.data, None, d.ctypes.data, stride), stride

        def pyramid(p):
            if isinstance(p, _lib.PyramidView):
                return p, _lib.DEVICE_PYRAMIDS
            v = _lib.PyramidView()
            v.nlevels = len(p[0])
            for l in range(v.nlevels):
                lv = np.ascontiguousarray(np.stack([np.asarray(q[l], np.uint8) for q in p]))
                keep.append(lv)
                v.level[l] = lv.ctypes.data; v.frame_stride[l] = lv.strides[0]; v.pitch[l] = lv.strides[1]
                v.height[l], v.width[l] = lv.shape[1], lv.shape[2]
            return v, 0

        lv, ls = side(lefts)
        rv, rs = side(rights)
        self._ensure(items, max(ls, rs))
        lp, lflag = pyramid(lpyr)
        rp, rflag = pyramid(rpyr)
        assert lflag == rflag, "both pyramids must live on the same side"
        sc = np.ascontiguousarray(scale_factors, np.float32); isc = np.ascontiguousarray(inv_scale_factors, np.float32)
        ur = np.zeros((items, ls), np.float32); dep = np.zeros((items, ls), np.float32); nm = np.zeros(items, np.int32)
        check(self._L.orbb200_compute_stereo_matches(self._h, items, C.byref(lv), C.byref(rv), C.byref(lp), C.byref(rp), sc.ctypes.data,
                                                     isc.ctypes.data, len(sc), float(mb), float(mbf), ur.ctypes.data, dep.ctypes.data,
                                                     nm.ctypes.data, lflag))
        return [(ur[i, :len(lefts[i][0])], dep[i, :len(lefts[i][0])]) for i in range(items)], nm

    # ---- SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (S/ORBmatcher.cc:661-827), scope row N3 ----
    def search_for_triangulation_batch(self, pairs, only_stereo=False, device_reps=0):
        """pairs: list of workloads.triangulation_pair()-layout dicts (k1 / k2: x, y, octave, angle, desc, has_mp, u_right,
        node, start, feat; F12 (9), epipole (2), scale_factors, level_sigma2).  Returns (nmatches (items,),
        [matches12 per item: index in key frame 2 or -1]); vMatchedPairs = [(i, m[i]) for i where m[i] >= 0]."""
        from ._lib import TriView
        items = len(pairs)
        a1, v1, s1 = _bow_side([p["k1"] for p in pairs])
        a2, v2, s2 = _bow_side([p["k2"] for p in pairs])
        self._ensure(items, max(s1, s2))

        def geo(key, stride):
            g = dict(x=_pack([p[key]["x"] for p in pairs], stride, np.float32), y=_pack([p[key]["y"] for p in pairs], stride, np.float32),
                     o=_pack([p[key]["octave"] for p in pairs], stride, np.int32), u=_pack([p[key]["u_right"] for p in pairs], stride, np.float32),
                     m=_pack([p[key]["has_mp"] for p in pairs], stride, np.uint8))
            return g, TriView(g["x"].ctypes.data, g["y"].ctypes.data, g["o"].ctypes.data, g["u"].ctypes.data, g["m"].ctypes.data)
        g1, t1 = geo("k1", s1)
        g2, t2 = geo("k2", s2)
        F = np.ascontiguousarray(np.stack([np.asarray(p["F12"], np.float32).reshape(9) for p in pairs]))
        ep = np.ascontiguousarray(np.stack([np.asarray(p["epipole"], np.float32).reshape(2) for p in pairs]))
        sf = np.ascontiguousarray(pairs[0]["scale_factors"], np.float32)
        ls = np.ascontiguousarray(pairs[0]["level_sigma2"], np.float32)
        m = np.full((items, s1), -1, np.int32)
        nm = np.zeros(items, np.int32)
        check(self._L.orbb200_search_for_triangulation(
            self._h, items, C.byref(v1), C.byref(t1), C.byref(v2), C.byref(t2), F.ctypes.data, ep.ctypes.data, sf.ctypes.data,
            ls.ctypes.data, len(sf), int(only_stereo), int(self.mbCheckOrientation), m.ctypes.data, nm.ctypes.data, 0))
        if device_reps:          # the same call on device-resident copies of the same data, CUDA events on the matcher's stream
            keep = []
            dv1, dv2 = _device_copy(v1, list(a1.values()), self.device, keep), _device_copy(v2, list(a2.values()), self.device, keep)
            dt1, dt2 = _device_copy(t1, list(g1.values()), self.device, keep), _device_copy(t2, list(g2.values()), self.device, keep)
            dF, dep = _device_array(F, self.device, keep), _device_array(ep, self.device, keep)
            dsf, dls = _device_array(sf, self.device, keep), _device_array(ls, self.device, keep)
            dm, dnm = _device_array(m, self.device, keep), _device_array(nm, self.device, keep)
            self.last_device_ms = _device_time(self._L.orbb200_matcher_stream(self._h), self.device, lambda: check(
                self._L.orbb200_search_for_triangulation(self._h, items, C.byref(dv1), C.byref(dt1), C.byref(dv2), C.byref(dt2), dF, dep,
                                                         dsf, dls, len(sf), int(only_stereo),
                                                         int(self.mbCheckOrientation), dm, dnm, 1)), device_reps)
            assert np.array_equal(keep[-1].cpu().numpy(), nm), "device-resident call differs from the host-buffer call"
        return nm, [m[i, :a1["n"][i]] for i in range(items)]

    # ---- SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (S/ORBmatcher.cc:1476-1603), scope row N2 ----
    def search_by_projection_keyframe_batch(self, cur_frames, kfs, th=10.0, orb_dist=100):
        """Relocalisation search.  cur_frames: list of Frame (mvpMapPoints: -1 = free, anything else = holds a map
        point; on return the accepted key-frame slot indices).  kfs: list of dicts with valid (1 = usable map
        point: not bad, not in sAlreadyFound), wpos (n,3), mp_desc (n,32), mf_max, mf_min (raw mfMaxDistance /
        mfMinDistance), kf_angle, Rcw (9), tcw (3), Ow (3), K (4), log_scale."""
        from ._lib import KeyFrameView
        items = len(cur_frames)
        keep = []
        fv, s = _frame_view(cur_frames, keep)
        ks = max(1, max(len(k["valid"]) for k in kfs))
        self._ensure(items, max(s, ks))
        bnd = cur_frames[0].bounds
        sf = cur_frames[0].mvScaleFactors
        assert sf is not None, "Frame.mvScaleFactors is required"
        kpmp = _pack([f.mvpMapPoints for f in cur_frames], s, np.int32)
        kn = np.array([len(k["valid"]) for k in kfs], np.int32)
        a = dict(va=_pack([(np.asarray(k["valid"]) == 1).astype(np.uint8) for k in kfs], ks, np.uint8),
                 wp=_pack([np.asarray(k["wpos"], np.float32).reshape(-1, 3) for k in kfs], ks, np.float32, (3,)),
                 md=_pack([k["mp_desc"] for k in kfs], ks, np.uint8, (32,)), mx=_pack([k["mf_max"] for k in kfs], ks, np.float32),
                 mn=_pack([k["mf_min"] for k in kfs], ks, np.float32), an=_pack([k["kf_angle"] for k in kfs], ks, np.float32))
        kv = KeyFrameView(kn.ctypes.data, a["va"].ctypes.data, a["wp"].ctypes.data, a["md"].ctypes.data, a["mx"].ctypes.data,
                          a["mn"].ctypes.data, a["an"].ctypes.data, ks)
        R = np.ascontiguousarray(np.stack([np.asarray(k["Rcw"], np.float32).reshape(9) for k in kfs]))
        t = np.ascontiguousarray(np.stack([np.asarray(k["tcw"], np.float32).reshape(3) for k in kfs]))
        Ow = np.ascontiguousarray(np.stack([np.asarray(k["Ow"], np.float32).reshape(3) for k in kfs]))
        K = np.ascontiguousarray(kfs[0]["K"], np.float32)
        nm = np.zeros(items, np.int32)
        check(self._L.orbb200_search_by_projection_keyframe(
            self._h, items, C.byref(fv), C.byref(kv), R.ctypes.data, t.ctypes.data, Ow.ctypes.data, K.ctypes.data,
            kpmp.ctypes.data, sf.ctypes.data, len(sf), float(kfs[0]["log_scale"]), bnd.ctypes.data, float(th), int(orb_dist),
            int(self.mbCheckOrientation), nm.ctypes.data, 0))
        for i, f in enumerate(cur_frames):
            f.mvpMapPoints[:] = kpmp[i, :f.N]
        return nm

    # ---- SearchByProjection(CurrentFrame, LastFrame, th, bMono) (S/ORBmatcher.cc:1332-1474), scope row N2 ----
    def search_by_projection_last_frame_batch(self, cur_frames, lasts, th=15.0, mode=0, mbf=40.0):
        """cur_frames: list of Frame (mvpMapPoints hold indices into the last frame's arrays).
        lasts: list of dicts with has_mp, outlier, wpos (n,3), mp_desc (n,32), mp_obs, last_octave, last_angle,
        Rcw (9), tcw (3), K (4) -- the fields of LastFrame / CurrentFrame.mTcw the reference reads.
        mode: 0 = bMono or neither forward nor backward, 1 = bForward, 2 = bBackward."""
        from ._lib import LastFrameView
        items = len(cur_frames)
        keep = []
        fv, s = _frame_view(cur_frames, keep)
        ls = max(1, max(len(l["has_mp"]) for l in lasts))
        self._ensure(items, max(s, ls))
        bnd = cur_frames[0].bounds
        sf = cur_frames[0].mvScaleFactors
        assert sf is not None, "Frame.mvScaleFactors is required"
        ur = _pack([f.mvuRight for f in cur_frames], s, np.float32)
        kpmp = _pack([f.mvpMapPoints for f in cur_frames], s, np.int32)
        kpobs = _pack([f.mvpMapPointObs for f in cur_frames], s, np.int32)
        ln = np.array([len(l["has_mp"]) for l in lasts], np.int32)
        a = dict(hm=_pack([l["has_mp"] for l in lasts], ls, np.uint8), ol=_pack([l["outlier"] for l in lasts], ls, np.uint8),
                 wp=_pack([np.asarray(l["wpos"], np.float32).reshape(-1, 3) for l in lasts], ls, np.float32, (3,)),
                 md=_pack([l["mp_desc"] for l in lasts], ls, np.uint8, (32,)), ob=_pack([l["mp_obs"] for l in lasts], ls, np.int32),
                 oc=_pack([l["last_octave"] for l in lasts], ls, np.int32), an=_pack([l["last_angle"] for l in lasts], ls, np.float32))
        lv = LastFrameView(ln.ctypes.data, a["hm"].ctypes.data, a["ol"].ctypes.data, a["wp"].ctypes.data, a["md"].ctypes.data,
                           a["ob"].ctypes.data, a["oc"].ctypes.data, a["an"].ctypes.data, ls)
        R = np.ascontiguousarray(np.stack([np.asarray(l["Rcw"], np.float32).reshape(9) for l in lasts]))
        t = np.ascontiguousarray(np.stack([np.asarray(l["tcw"], np.float32).reshape(3) for l in lasts]))
        K = np.ascontiguousarray(lasts[0]["K"], np.float32)
        nm = np.zeros(items, np.int32)
        check(self._L.orbb200_search_by_projection_last_frame(
            self._h, items, C.byref(fv), ur.ctypes.data, C.byref(lv), R.ctypes.data, t.ctypes.data, K.ctypes.data, float(mbf),
            kpmp.ctypes.data, kpobs.ctypes.data, sf.ctypes.data, len(sf), bnd.ctypes.data, float(th), int(mode),
            int(self.mbCheckOrientation), nm.ctypes.data, 0))
        for i, f in enumerate(cur_frames):
            f.mvpMapPoints[:] = kpmp[i, :f.N]
        return nm


class Vocabulary:
    """A DBoW2 vocabulary tree on the device (include/orb_b200.h, orbb200_vocabulary_create).  voc: dict with child_start,
    children, desc (n_nodes, 32), word_id, weight (float64), L -- workloads.synthetic_vocabulary() layout; a vocabulary
    read from ORBvoc.txt flattens the same way (node id = line number, children in file order)."""

    def __init__(self, voc, device=0):
        self._L = _lib.load()
        self._h = _lib.vp()
        self.device = device
        cs = np.ascontiguousarray(voc["child_start"], np.int32); ch = np.ascontiguousarray(voc["children"], np.int32)
        d = np.ascontiguousarray(voc["desc"], np.uint8); wi = np.ascontiguousarray(voc["word_id"], np.int32)
        wt = np.ascontiguousarray(voc["weight"], np.float64)
        check(self._L.orbb200_vocabulary_create(device, len(wi), int(voc["L"]), cs.ctypes.data, ch.ctypes.data, d.ctypes.data,
                                                wi.ctypes.data, wt.ctypes.data, C.byref(self._h)))

    def close(self):
        if self._h:
            self._L.orbb200_vocabulary_destroy(self._h)
            self._h = _lib.vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bow_transform_batch(matcher, vocabulary, descs, levelsup=4):
    """Frame::ComputeBoW for a list of (n_i, 32) uint8 descriptor arrays.  Returns a list of dicts
    (word, value, node, start, feat) -- BowVector and flattened FeatureVector per frame."""
    items = len(descs)
    n = np.array([len(d) for d in descs], np.int32)
    s = max(1, int(n.max()))
    matcher._ensure(items, s)
    dd = _pack([np.asarray(d, np.uint8).reshape(-1, 32) for d in descs], s, np.uint8, (32,))
    bn, fn = np.zeros(items, np.int32), np.zeros(items, np.int32)
    bw, fnode, ffeat = (np.zeros((items, s), np.uint32) for _ in range(3))
    bv = np.zeros((items, s), np.float64)
    fstart = np.zeros((items, s + 1), np.int32)
    check(matcher._L.orbb200_bow_transform(matcher._h, vocabulary._h, items, n.ctypes.data, dd.ctypes.data, s, int(levelsup), bn.ctypes.data,
                                           bw.ctypes.data, bv.ctypes.data, fn.ctypes.data, fnode.ctypes.data, fstart.ctypes.data,
                                           ffeat.ctypes.data, 0))
    return [dict(word=bw[i, :bn[i]], value=bv[i, :bn[i]], node=fnode[i, :fn[i]], start=fstart[i, :fn[i] + 1], feat=ffeat[i, :fstart[i, fn[i]]])
            for i in range(items)]
