"""ctypes binding of oracle/_ref/libref_orb.so: the reference's OWN unmodified ORBextractor.cc
compiled for x86 against mini-cv (see oracle/ref_harness.cc).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from oracle_lib import KP_DTYPE, ORACLE_DIR, build_oracle

REF_SO = os.path.join(ORACLE_DIR, "_ref", "libref_orb.so")
_u8p = C.POINTER(C.c_uint8)
_lib = None


def available():
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(REF_SO)
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
        _lib = L
    return _lib


class RefExtractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
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
