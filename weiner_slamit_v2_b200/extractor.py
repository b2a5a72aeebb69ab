"""Host-side mirror of ORB_SLAM2::ORBextractor (I/ORBextractor.h:45-111) over the C ABI.

Same constructor arguments, same getters, same call semantics as the reference class:
`extractor(image)` returns the keypoints (cv::KeyPoint records, reference order) and the
n x 32 descriptor matrix; an empty image returns without touching the outputs
(S/ORBextractor.cc:1068); a frame with no keypoints yields an empty descriptor matrix (:1095).
`extract_batch` is the B200 addition: many independent frames per call.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, addr, check


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width=640, height=480,
                 max_batch=1, device=0, blur_taps=0):
        self._L = _lib.load()
        self._h = _lib.vp()
        check(self._L.orbb200_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width, height,
                                               max_batch, device, blur_taps, C.byref(self._h)))
        self.nfeatures, self.nlevels, self.width, self.height = nfeatures, nlevels, width, height
        self.max_batch, self.device = max_batch, device
        self._scale_factor = float(np.float32(scaleFactor))
        self.max_keypoints = self._L.orbb200_extractor_max_keypoints(self._h)
        n = nlevels
        self._scale, self._inv_scale, self._sigma2, self._inv_sigma2 = (np.zeros(n, np.float32) for _ in range(4))
        self._per_level = np.zeros(n, np.int32)
        self._umax = np.zeros(16, np.int32)
        check(self._L.orbb200_extractor_tables(self._h, *(a.ctypes.data for a in (
            self._scale, self._inv_scale, self._sigma2, self._inv_sigma2, self._per_level, self._umax))))

    def close(self):
        if getattr(self, "_h", None):
            self._L.orbb200_extractor_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference getters (I/ORBextractor.h:63-83) ----
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self._scale_factor
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv_scale.copy()
    def GetScaleSigmaSquares(self): return self._sigma2.copy()
    def GetInverseScaleSigmaSquares(self): return self._inv_sigma2.copy()
    @property
    def mnFeaturesPerLevel(self): return self._per_level.copy()
    @property
    def umax(self): return self._umax.copy()

    # ---- operator() ----
    def __call__(self, image, mask=None):
        """(keypoints, descriptors) for one H x W uint8 frame; mask is ignored like the reference."""
        if image is None or image.size == 0:
            return None
        kps, desc, counts = self.extract_batch(np.asarray(image)[None])
        n = int(counts[0])
        return kps[0, :n].copy(), desc[0, :n].copy()

    def extract_batch(self, images):
        """images: (B, H, W) uint8 numpy array in host memory.  Returns (keypoints (B,cap) KP_DTYPE,
        descriptors (B,cap,32) uint8, counts (B,) int32)."""
        images = np.asarray(images)
        if images.dtype != np.uint8 or images.ndim != 3:
            raise ValueError("expected a (B, H, W) uint8 array")          # CV_8UC1 assert (:1075)
        b, hgt, w = images.shape
        if (hgt, w) != (self.height, self.width):
            raise ValueError("frame is %dx%d, handle was created for %dx%d" % (w, hgt, self.width, self.height))
        if images.strides[2] != 1 or images.strides[0] < 0 or images.strides[1] < w:
            images = np.ascontiguousarray(images)
        cap = self.max_keypoints
        kps = np.zeros((b, cap), KP_DTYPE)
        desc = np.zeros((b, cap, 32), np.uint8)
        counts = np.zeros(b, np.int32)
        check(self._L.orbb200_extract_host(self._h, images.ctypes.data, b, images.strides[1], images.strides[0],
                                           kps.ctypes.data, desc.ctypes.data, counts.ctypes.data, cap))
        return kps, desc, counts

    def extract_host_async(self, images, batch, stride, frame_stride, kps, desc, counts, cap):
        """Queue one batch from (page-locked) host memory; arguments are arrays, tensors or raw addresses.
        wait() makes the results visible.  One call in flight per handle (see StreamingExtractor)."""
        check(self._L.orbb200_extract_host_async(self._h, addr(images), batch, stride, frame_stride, addr(kps),
                                                 addr(desc), addr(counts), cap))

    def wait(self):
        check(self._L.orbb200_extract_host_wait(self._h))

    def extract_device(self, d_images, batch, stride, frame_stride, d_kps=None, d_desc=None, d_counts=None, cap=0):
        """Asynchronous: frames already in device memory (torch tensors or raw addresses)."""
        check(self._L.orbb200_extract_device(self._h, addr(d_images), batch, stride, frame_stride, addr(d_kps),
                                             addr(d_desc), addr(d_counts), cap))

    def sync(self):
        check(self._L.orbb200_extractor_sync(self._h))

    def device_outputs(self):
        k, d, c, cap = _lib.vp(), _lib.vp(), _lib.vp(), C.c_int()
        check(self._L.orbb200_extractor_outputs(self._h, C.byref(k), C.byref(d), C.byref(c), C.byref(cap)))
        return k.value, d.value, c.value, cap.value

    @property
    def stream(self): return self._L.orbb200_extractor_stream(self._h)
    @property
    def last_launches(self): return self._L.orbb200_extractor_last_launches(self._h)

    def set_profiling(self, on=True):
        check(self._L.orbb200_extractor_set_profiling(self._h, int(on)))

    def stage_ms(self):
        """Device ms of the last call per stage: pyramid, fast, quadtree, blur, describe."""
        ms = np.zeros(5, np.float32)
        check(self._L.orbb200_extractor_stage_ms(self._h, ms.ctypes.data))
        return ms

    def pyramid_view(self):
        """mvImagePyramid of the last batch where it lies in device memory (for compute_stereo_matches)."""
        v = _lib.PyramidView()
        check(self._L.orbb200_extractor_pyramid_view(self._h, C.byref(v)))
        return v

    # ---- stage read-back (parity tests; mvImagePyramid) ----
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(self._L.orbb200_extractor_level_size(self._h, level, C.byref(w), C.byref(h)))
        return h.value, w.value

    def get_level(self, frame, level, blurred=False):
        h, w = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        check(self._L.orbb200_extractor_get_level(self._h, frame, level, int(blurred), out.ctypes.data, w))
        return out

    def _packed(self, fn, frame, level, cap):
        x, y, s = (np.zeros(cap, np.int32) for _ in range(3))
        n = C.c_int()
        check(fn(self._h, frame, level, x.ctypes.data, y.ctypes.data, s.ctypes.data, cap, C.byref(n)))
        return x[:n.value], y[:n.value], s[:n.value]

    def get_candidates(self, frame, level):
        h, w = self.level_size(level)
        return self._packed(self._L.orbb200_extractor_get_candidates, frame, level, (w // 2 + 2) * (h // 2 + 2))

    def get_level_keypoints(self, frame, level):
        return self._packed(self._L.orbb200_extractor_get_level_keypoints, frame, level, self.max_keypoints)


class StreamingExtractor:
    """A stream of host batches through `depth` extractor handles used in turn, so that the upload of batch
    k+1 and the download of batch k-1 run beside the kernels of batch k (each handle has its own streams and
    slabs; include/orb_b200.h, orbb200_extract_host_async).  submit() returns the index of the slot it used;
    that slot's previous batch is complete when submit() returns, and every batch is complete after drain().
    Host buffers should be page-locked."""

    def __init__(self, *args, depth=3, **kw):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.handles = [ORBextractor(*args, **kw) for _ in range(depth)]
        self._next = 0

    @property
    def max_keypoints(self): return self.handles[0].max_keypoints

    def submit(self, images, batch, stride, frame_stride, kps, desc, counts, cap):
        slot = self._next
        self._next = (slot + 1) % len(self.handles)
        self.handles[slot].extract_host_async(images, batch, stride, frame_stride, kps, desc, counts, cap)
        return slot

    def wait(self, slot):
        self.handles[slot].wait()

    def drain(self):
        for h in self.handles:
            h.wait()

    @property
    def last_launches(self): return sum(h.last_launches for h in self.handles)

    def close(self):
        for h in self.handles:
            h.close()
