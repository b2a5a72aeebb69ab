"""weiner_slamit_v2_b200 -- the ORB-SLAM2 front-end hot path of serviceberry3/weiner_slamit_v2
(ORBextractor::operator() and ORBmatcher's Hamming search) on B200, behind a C ABI
(include/orb_b200.h, built into liborb_b200.so by csrc/build.sh).  No CPU fallback."""
from ._lib import KP_DTYPE, OrbB200Error, load  # noqa: F401
from .extractor import ORBextractor, StreamingExtractor  # noqa: F401
from . import frames  # noqa: F401
