"""Device-resident front end: ORBextractor -> Frame glue (undistort, SoA, grid) -> ORBmatcher without a
host round trip.  This is what the reference does per frame pair in
Tracking::MonocularInitialization (S/Tracking.cc:750-849): two Frame constructors (extract, undistort, grid)
followed by ORBmatcher(0.9, true).SearchForInitialization(..., 100) -- here for a batch of pairs at once.
torch is used for the device buffers only."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FrameView, check
from .extractor import ORBextractor

# the camera the reference hard-codes (S/Tracking.cc:77-80, 101-111)
REFERENCE_K = np.array([526.69, 540.36, 313.07, 238.39], np.float32)
REFERENCE_DIST = np.array([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], np.float32)


class InitializationPipeline:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, width=640, height=480,
                 max_pairs=128, K=REFERENCE_K, dist=REFERENCE_DIST, nnratio=0.9, check_orientation=True, window=100,
                 device=0):
        self.L = _lib.load()
        self.dev = torch.device("cuda", device)
        self.ex = ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height, max_pairs, device)
        self.cap = self.ex.max_keypoints
        self.m = _lib.vp()
        check(self.L.orbb200_matcher_create(max_pairs, self.cap, device, C.byref(self.m)))
        self.K = np.ascontiguousarray(K, np.float32)
        self.dist = np.ascontiguousarray(dist, np.float32)
        self.nnratio, self.check_ori, self.window = float(nnratio), int(check_orientation), int(window)
        self.bounds = np.zeros(4, np.float32)
        check(self.L.orbb200_image_bounds(self.m, width, height, self.K.ctypes.data, self.dist.ctypes.data, self.bounds.ctypes.data))
        p, c = max_pairs, self.cap
        def buf(*shape, dtype=torch.float32):
            return torch.empty(shape, dtype=dtype, device=self.dev)
        self.side = []
        for _ in range(2):
            self.side.append(dict(kps=buf(p, c, 28, dtype=torch.uint8), desc=buf(p, c, 32, dtype=torch.uint8),
                                  n=buf(p, dtype=torch.int32), x=buf(p, c), y=buf(p, c), oct=buf(p, c, dtype=torch.int32),
                                  ang=buf(p, c)))
        self.prev = buf(p, c, 2)
        self.m12 = buf(p, c, dtype=torch.int32)
        self.nm = buf(p, dtype=torch.int32)

    def close(self):
        if self.m:
            self.L.orbb200_matcher_destroy(self.m)
            self.m = None
        self.ex.close()

    def _view(self, s):
        return FrameView(s["n"].data_ptr(), s["x"].data_ptr(), s["y"].data_ptr(), s["oct"].data_ptr(), s["ang"].data_ptr(),
                         s["desc"].data_ptr(), self.cap)

    def run(self, d_frames1, d_frames2, pairs):
        """d_frames1/2: (pairs, H, W) uint8 device tensors.  Everything is queued asynchronously; returns the
        device tensors (nmatches, matches12, prev_matched, side buffers)."""
        h, w = d_frames1.shape[1:]
        for s, fr in zip(self.side, (d_frames1, d_frames2)):
            self.ex.extract_device(fr, pairs, w, w * h, s["kps"], s["desc"], s["n"], self.cap)
        check(self.L.orbb200_matcher_wait_extractor(self.m, self.ex._h))
        for s in self.side:
            check(self.L.orbb200_frames_from_keypoints(self.m, s["kps"].data_ptr(), s["n"].data_ptr(), pairs, self.cap,
                                                       self.K.ctypes.data, self.dist.ctypes.data, s["x"].data_ptr(),
                                                       s["y"].data_ptr(), s["oct"].data_ptr(), s["ang"].data_ptr()))
        st = torch.cuda.ExternalStream(self.L.orbb200_matcher_stream(self.m), device=self.dev)
        with torch.cuda.stream(st):     # vbPrevMatched = F1.mvKeysUn[i].pt (S/Tracking.cc:776-778)
            self.prev[:pairs] = torch.stack([self.side[0]["x"][:pairs], self.side[0]["y"][:pairs]], dim=-1)
        v1, v2 = self._view(self.side[0]), self._view(self.side[1])
        check(self.L.orbb200_search_for_initialization(self.m, pairs, C.byref(v1), C.byref(v2), self.bounds.ctypes.data,
                                                       self.nnratio, self.check_ori, self.window, self.prev.data_ptr(),
                                                       self.m12.data_ptr(), self.nm.data_ptr(), 1))
        return self.nm, self.m12, self.prev

    def sync(self):
        self.ex.sync()
        check(self.L.orbb200_matcher_sync(self.m))


class StereoPipeline:
    """The stereo Frame constructor (S/Frame.cc:60-116) for a batch of rectified pairs without leaving the device: the
    two extractors run on their own handles and streams (the reference starts two threads, :93-96), the keypoints are
    turned into structure-of-arrays views, and Frame::ComputeStereoMatches (:591-763) reads both pyramids where the
    extractors left them."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, width=640, height=480,
                 max_pairs=128, mb=0.1, mbf=40.0, device=0):
        self.L = _lib.load()
        self.dev = torch.device("cuda", device)
        self.exl = ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height, max_pairs, device)
        self.exr = ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th, width, height, max_pairs, device)
        self.cap = self.exl.max_keypoints
        self.m = _lib.vp()
        check(self.L.orbb200_matcher_create(max_pairs, self.cap, device, C.byref(self.m)))
        self.mb, self.mbf = float(mb), float(mbf)
        self.scale = self.exl.GetScaleFactors(); self.inv_scale = self.exl.GetInverseScaleFactors()
        self.K = np.array([1, 1, 0, 0], np.float32); self.dist = np.zeros(5, np.float32)     # mvKeys, not mvKeysUn
        p, c = max_pairs, self.cap
        def buf(*shape, dtype=torch.float32):
            return torch.empty(shape, dtype=dtype, device=self.dev)
        self.side = [dict(kps=buf(p, c, 28, dtype=torch.uint8), desc=buf(p, c, 32, dtype=torch.uint8), n=buf(p, dtype=torch.int32),
                          x=buf(p, c), y=buf(p, c), oct=buf(p, c, dtype=torch.int32), ang=buf(p, c)) for _ in range(2)]
        self.u_right = buf(p, c); self.depth = buf(p, c); self.nm = buf(p, dtype=torch.int32)

    def close(self):
        if self.m:
            self.L.orbb200_matcher_destroy(self.m)
            self.m = None
        self.exl.close(); self.exr.close()

    def run(self, d_left, d_right, pairs):
        """d_left / d_right: (pairs, H, W) uint8 device tensors.  Asynchronous; returns the device tensors
        (nmatches, mvuRight, mvDepth); side[0] holds the left keypoints and descriptors."""
        h, w = d_left.shape[1:]
        for ex, s, fr in zip((self.exl, self.exr), self.side, (d_left, d_right)):
            ex.extract_device(fr, pairs, w, w * h, s["kps"], s["desc"], s["n"], self.cap)
            check(self.L.orbb200_matcher_wait_extractor(self.m, ex._h))
        views = []
        for s in self.side:
            check(self.L.orbb200_frames_from_keypoints(self.m, s["kps"].data_ptr(), s["n"].data_ptr(), pairs, self.cap,
                                                       self.K.ctypes.data, self.dist.ctypes.data, s["x"].data_ptr(),
                                                       s["y"].data_ptr(), s["oct"].data_ptr(), s["ang"].data_ptr()))
            views.append(FrameView(s["n"].data_ptr(), s["x"].data_ptr(), s["y"].data_ptr(), s["oct"].data_ptr(), s["ang"].data_ptr(),
                                   s["desc"].data_ptr(), self.cap))
        lp, rp = self.exl.pyramid_view(), self.exr.pyramid_view()
        check(self.L.orbb200_compute_stereo_matches(self.m, pairs, C.byref(views[0]), C.byref(views[1]), C.byref(lp), C.byref(rp),
                                                    self.scale.ctypes.data, self.inv_scale.ctypes.data, len(self.scale), self.mb, self.mbf,
                                                    self.u_right.data_ptr(), self.depth.data_ptr(), self.nm.data_ptr(),
                                                    _lib.DEVICE_VIEWS | _lib.DEVICE_PYRAMIDS))
        return self.nm, self.u_right, self.depth

    def sync(self):
        self.exl.sync(); self.exr.sync()
        check(self.L.orbb200_matcher_sync(self.m))
