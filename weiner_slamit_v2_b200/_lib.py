"""ctypes loader for liborb_b200.so (the C ABI declared in include/orb_b200.h).

There is no fallback of any kind: if the CUDA library has not been built, or a compute entry
point fails, an exception is raised.  Nothing under oracle/ is ever imported from here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborb_b200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p


class OrbB200Error(RuntimeError):
    pass


class FrameView(C.Structure):
    _fields_ = [("n", vp), ("x", vp), ("y", vp), ("octave", vp), ("angle", vp), ("desc", vp), ("stride", C.c_int)]


class MapPointView(C.Structure):
    _fields_ = [("n", vp), ("in_view", vp), ("bad", vp), ("proj_x", vp), ("proj_y", vp), ("proj_xr", vp),
                ("level", vp), ("view_cos", vp), ("desc", vp), ("obs", vp), ("stride", C.c_int)]


class LastFrameView(C.Structure):
    _fields_ = [("n", vp), ("has_mp", vp), ("outlier", vp), ("world_pos", vp), ("mp_desc", vp), ("mp_obs", vp),
                ("octave", vp), ("angle", vp), ("stride", C.c_int)]


class KeyFrameView(C.Structure):
    _fields_ = [("n", vp), ("valid", vp), ("world_pos", vp), ("mp_desc", vp), ("max_distance", vp), ("min_distance", vp),
                ("angle", vp), ("stride", C.c_int)]


class BowView(C.Structure):
    _fields_ = [("n", vp), ("desc", vp), ("angle", vp), ("valid", vp), ("n_nodes", vp), ("node_id", vp), ("node_start", vp),
                ("feat", vp), ("stride", C.c_int), ("node_stride", C.c_int)]


class TriView(C.Structure):
    _fields_ = [("x", vp), ("y", vp), ("octave", vp), ("u_right", vp), ("has_mp", vp)]


class FusePointsView(C.Structure):
    _fields_ = [("n", vp), ("valid", vp), ("world_pos", vp), ("normal", vp), ("mp_desc", vp), ("max_distance", vp), ("min_distance", vp),
                ("stride", C.c_int)]


MAX_LEVELS = 16


class PyramidView(C.Structure):
    _fields_ = [("nlevels", C.c_int), ("level", vp * MAX_LEVELS), ("frame_stride", C.c_size_t * MAX_LEVELS),
                ("pitch", C.c_int32 * MAX_LEVELS), ("width", C.c_int32 * MAX_LEVELS), ("height", C.c_int32 * MAX_LEVELS)]


DEVICE_VIEWS, DEVICE_PYRAMIDS = 1, 2

# name -> (restype, argtypes); every symbol include/orb_b200.h declares
SIGNATURES = {
    "orbb200_last_error": (C.c_char_p, []),
    "orbb200_device_count": (C.c_int, []),
    "orbb200_version": (C.c_char_p, []),
    "orbb200_extractor_create": (C.c_int, [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.POINTER(vp)]),
    "orbb200_extractor_destroy": (None, [vp]),
    "orbb200_extractor_tables": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "orbb200_extractor_levels": (C.c_int, [vp]),
    "orbb200_extractor_max_keypoints": (C.c_int, [vp]),
    "orbb200_extractor_level_size": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "orbb200_extract_host": (C.c_int, [vp, vp, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int]),
    "orbb200_extract_host_async": (C.c_int, [vp, vp, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int]),
    "orbb200_extract_host_wait": (C.c_int, [vp]),
    "orbb200_extract_device": (C.c_int, [vp, vp, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int]),
    "orbb200_extractor_sync": (C.c_int, [vp]),
    "orbb200_extractor_outputs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int)]),
    "orbb200_extractor_stream": (vp, [vp]),
    "orbb200_extractor_last_launches": (C.c_int, [vp]),
    "orbb200_extractor_set_profiling": (C.c_int, [vp, C.c_int]),
    "orbb200_extractor_stage_ms": (C.c_int, [vp, vp]),
    "orbb200_extractor_get_level": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]),
    "orbb200_extractor_get_candidates": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]),
    "orbb200_extractor_get_level_keypoints": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]),
    "orbb200_matcher_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "orbb200_matcher_destroy": (None, [vp]),
    "orbb200_matcher_stream": (vp, [vp]),
    "orbb200_matcher_sync": (C.c_int, [vp]),
    "orbb200_matcher_last_launches": (C.c_int, [vp]),
    "orbb200_matcher_device": (C.c_int, [vp]),
    "orbb200_extractor_device": (C.c_int, [vp]),
    "orbb200_extractor_debug_set_capacity": (C.c_int, [vp, C.c_int]),
    "orbb200_descriptor_distance": (C.c_int, [vp, vp, vp, C.c_int, vp]),
    "orbb200_search_for_initialization": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), vp,
                                                    C.c_float, C.c_int, C.c_int, vp, vp, vp, C.c_int]),
    "orbb200_search_by_projection": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), vp, C.POINTER(MapPointView), vp, vp,
                                               vp, C.c_int, vp, C.c_float, C.c_float, vp, C.c_int]),
    "orbb200_search_by_projection_last_frame": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), vp, C.POINTER(LastFrameView), vp, vp,
                                                          vp, C.c_float, vp, vp, vp, C.c_int, vp, C.c_float, C.c_int, C.c_int,
                                                          vp, C.c_int]),
    "orbb200_search_by_projection_keyframe": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), C.POINTER(KeyFrameView), vp, vp, vp, vp,
                                                        vp, vp, C.c_int, C.c_float, vp, C.c_float, C.c_int, C.c_int, vp, C.c_int]),
    "orbb200_search_by_bow": (C.c_int, [vp, C.c_int, C.POINTER(BowView), C.POINTER(BowView), C.c_float, C.c_int, vp, vp, C.c_int]),
    "orbb200_search_by_bow_keyframes": (C.c_int, [vp, C.c_int, C.POINTER(BowView), C.POINTER(BowView), C.c_float, C.c_int, vp, vp, C.c_int]),
    "orbb200_search_for_triangulation": (C.c_int, [vp, C.c_int, C.POINTER(BowView), C.POINTER(TriView), C.POINTER(BowView),
                                                   C.POINTER(TriView), vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]),
    "orbb200_distinctive_descriptors": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp, vp, C.c_int]),
    "orbb200_fuse_search": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), vp, C.POINTER(FusePointsView), vp, vp, vp, vp, C.c_float, vp, vp,
                                      C.c_int, C.c_float, vp, C.c_float, C.c_int, vp, vp, vp, vp, C.c_int]),
    "orbb200_search_by_projection_sim3": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), C.POINTER(FusePointsView), vp, vp, vp, vp, vp,
                                                    C.c_int, C.c_float, vp, C.c_int, vp, vp, C.c_int]),
    "orbb200_vocabulary_create": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "orbb200_vocabulary_destroy": (None, [vp]),
    "orbb200_bow_transform": (C.c_int, [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int]),
    "orbb200_frames_from_keypoints": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]),
    "orbb200_undistort_points": (C.c_int, [vp, vp, vp, C.c_int, vp, vp]),
    "orbb200_image_bounds": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
    "orbb200_matcher_wait_extractor": (C.c_int, [vp, vp]),
    "orbb200_extractor_pyramid_view": (C.c_int, [vp, C.POINTER(PyramidView)]),
    "orbb200_compute_stereo_matches": (C.c_int, [vp, C.c_int, C.POINTER(FrameView), C.POINTER(FrameView), C.POINTER(PyramidView),
                                                 C.POINTER(PyramidView), vp, vp, C.c_int, C.c_float, C.c_float, vp, vp, vp, C.c_int]),
}

_lib = None


def load():
    """Load liborb_b200.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrbB200Error("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise OrbB200Error("orb_b200 error %d: %s" % (rc, load().orbb200_last_error().decode()))


def addr(a):
    """Raw address of a numpy array, a torch tensor, an int, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()     # torch.Tensor
