"""Generates tests/golden/*.npz from the reference's OWN ORBextractor.cc compiled into oracle/_ref
(run in the build container; the GPU box only reads the committed vectors).

Each file holds, for a few seeded synthetic frames, the reference's keypoints (cv::KeyPoint records)
and descriptors, plus a digest of every pyramid level.  Frames are regenerated from their seed by
weiner_slamit_v2_b200.frames, so the vectors stay small."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_lib as R  # noqa: E402
from weiner_slamit_v2_b200.frames import low_contrast_frame, synthetic_frame  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
assert R.available(), "build oracle/_ref first (make -C oracle ref)"


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def dump(name, params, frames, meta):
    ex = R.RefExtractor(*params)
    out = {"params": np.array(params, np.float64), "meta": np.array(meta)}
    for i, img in enumerate(frames):
        k, d = ex(img)
        out["kps_%d" % i] = k
        out["desc_%d" % i] = d
        out["pyr_sha_%d" % i] = np.stack([digest(ex.level_pixels(l)) for l in range(params[2])])
        out["frame_sha_%d" % i] = digest(img)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, [len(out["kps_%d" % i]) for i in range(len(frames))])


dump("ref_extract_640x480_n1000.npz", (1000, 1.2, 8, 20, 7), [synthetic_frame(i) for i in range(3)],
     ["synthetic_frame", "640", "480", "0", "1", "2"])
dump("ref_extract_1280x720_n2000.npz", (2000, 1.2, 8, 20, 7), [synthetic_frame(i, 1280, 720) for i in range(1)],
     ["synthetic_frame", "1280", "720", "0"])
dump("ref_extract_lowcontrast.npz", (1000, 1.2, 8, 20, 7), [low_contrast_frame(i) for i in range(2)],
     ["low_contrast_frame", "640", "480", "0", "1"])
dump("ref_extract_752x480_n500_s1.5_l5.npz", (500, 1.5, 5, 25, 9), [synthetic_frame(40 + i, 752, 480) for i in range(2)],
     ["synthetic_frame", "752", "480", "40", "41"])
