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

# ---- matcher vectors from the reference's own ORBmatcher.cc (oracle/_ref/libref_matcher.so) ----
assert R.matcher_available(), "build oracle/_ref/libref_matcher.so first (make -C oracle refm)"
from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, init_pair, projection_frame  # noqa: E402

out = {}
cfgs = [(400, 1000, 0, 100), (401, 1000, 1, 1000), (402, 600, 0, 30), (403, 250, 1, 50)]
for i, (idx, n, brute, window) in enumerate(cfgs):
    p = init_pair(idx, n=n, brute_force=bool(brute))
    r = R.ref_search_for_initialization(p[0], p[1], p[2], p[3], p[4], (0, 0, 640, 480), 0.9, True, window)
    out["cfg_%d" % i] = np.array([idx, n, brute, window]); out["n_%d" % i] = r[0]; out["m12_%d" % i] = r[1]; out["prev_%d" % i] = r[2]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_init.npz"), **out)
print("ref_match_init.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
cfgs = [(500, 2000, 10000, 1.0), (501, 2000, 6000, 3.0), (502, 800, 3000, 5.0)]
for i, (idx, nk, nmp, th) in enumerate(cfgs):
    kp, kd, mp = projection_frame(idx, nk, nmp)
    r = R.ref_search_by_projection(mp, kp, kd, SCALE_FACTORS_8, (0, 0, 1280, 720), 0.8, th)
    out["cfg_%d" % i] = np.array([idx, nk, nmp, th]); out["n_%d" % i] = r[0]; out["kpmp_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_proj.npz"), **out)
print("ref_match_proj.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
from weiner_slamit_v2_b200.workloads import motion_frame  # noqa: E402
cfgs = [(700, 15.0, 1), (701, 7.0, 1), (702, 15.0, 0)]
for i, (idx, th, ori) in enumerate(cfgs):
    w = motion_frame(idx)
    r = R.ref_search_by_projection_last_frame(w, SCALE_FACTORS_8, (-13.7, -9.2, 661.3, 492.8), th, bool(ori))
    out["cfg_%d" % i] = np.array([idx, th, ori]); out["n_%d" % i] = r[0]; out["kpmp_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_lastframe.npz"), **out)
print("ref_match_lastframe.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
from weiner_slamit_v2_b200.workloads import relocalisation_frame  # noqa: E402
cfgs = [(800, 10.0, 100, 1), (801, 3.0, 64, 1), (802, 10.0, 100, 0)]
for i, (idx, th, od, ori) in enumerate(cfgs):
    w = relocalisation_frame(idx)
    r = R.ref_search_by_projection_keyframe(w, SCALE_FACTORS_8, (-13.7, -9.2, 661.3, 492.8), th, od, bool(ori))
    out["cfg_%d" % i] = np.array([idx, th, od, ori]); out["n_%d" % i] = r[0]; out["kpmp_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_keyframe.npz"), **out)
print("ref_match_keyframe.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
from weiner_slamit_v2_b200.workloads import bow_pair  # noqa: E402
cfgs = [(850, 2000, 2000, 100, 0.7, 1), (851, 1500, 1800, 60, 0.9, 1), (852, 800, 700, 400, 0.7, 0)]
for i, (idx, nk, nf, nn, ratio, ori) in enumerate(cfgs):
    r = R.ref_search_by_bow(bow_pair(idx, nk, nf, nn), ratio, bool(ori))
    out["cfg_%d" % i] = np.array([idx, nk, nf, nn, ratio, ori]); out["n_%d" % i] = r[0]; out["m_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_bow.npz"), **out)
print("ref_match_bow.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
cfgs = [(860, 2000, 2000, 100, 0.75, 1), (861, 1500, 1800, 60, 0.9, 1), (862, 800, 700, 400, 0.75, 0)]
for i, (idx, nk, nf, nn, ratio, ori) in enumerate(cfgs):
    r = R.ref_search_by_bow_keyframes(bow_pair(idx, nk, nf, nn), ratio, bool(ori))
    out["cfg_%d" % i] = np.array([idx, nk, nf, nn, ratio, ori]); out["n_%d" % i] = r[0]; out["m_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_bowkf.npz"), **out)
print("ref_match_bowkf.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
from weiner_slamit_v2_b200.workloads import triangulation_pair  # noqa: E402
cfgs = [(870, 2000, 2000, 100, 0.0, 0, 0, 1), (871, 1500, 1800, 60, 0.3, 0, 1, 1), (872, 2000, 2000, 100, 0.0, 1, 0, 1), (873, 800, 700, 400, 0.2, 1, 0, 0)]
for i, (idx, na, nb, nn, sfr, fwd, st, ori) in enumerate(cfgs):
    r = R.ref_search_for_triangulation(triangulation_pair(idx, na, nb, nn, stereo_fraction=sfr, forward=bool(fwd)), bool(st), bool(ori))
    out["cfg_%d" % i] = np.array([idx, na, nb, nn, sfr, fwd, st, ori]); out["n_%d" % i] = r[0]; out["m_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_triangulation.npz"), **out)
print("ref_match_triangulation.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
from weiner_slamit_v2_b200.workloads import fuse_frame  # noqa: E402
cfgs = [(880, 3000, 2000, 3.0), (881, 2000, 1500, 2.5), (882, 800, 900, 10.0)]
for i, (idx, nmp, nkp, th) in enumerate(cfgs):
    r = R.ref_fuse_search(fuse_frame(idx, nmp, nkp), (-13.7, -9.2, 661.3, 492.8), th)
    out["cfg_%d" % i] = np.array([idx, nmp, nkp, th]); out["best_%d" % i] = r
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_fuse.npz"), **out)
print("ref_match_fuse.npz", [int((out["best_%d" % i] >= 0).sum()) for i in range(len(cfgs))])

out = {}
from weiner_slamit_v2_b200.workloads import sim3_pair  # noqa: E402
cfgs = [(883, 3000, 2000, 3.0), (884, 1500, 1500, 4.0)]
for i, (idx, nmp, nkp, th) in enumerate(cfgs):
    out["fcfg_%d" % i] = np.array([idx, nmp, nkp, th]); out["fbest_%d" % i] = R.ref_fuse_search_sim3(fuse_frame(idx, nmp, nkp), (-13.7, -9.2, 661.3, 492.8), th)
out["fcount"] = len(cfgs)
cfgs = [(890, 1500, 1500, 7.5), (891, 2000, 1700, 10.0)]
for i, (idx, na, nb, th) in enumerate(cfgs):
    r = R.ref_search_by_sim3(sim3_pair(idx, na, nb), (-13.7, -9.2, 661.3, 492.8), th)
    out["scfg_%d" % i] = np.array([idx, na, nb, th]); out["sn_%d" % i] = r[0]; out["sm_%d" % i] = r[1]
out["scount"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_sim3.npz"), **out)
print("ref_match_sim3.npz", [int((out["fbest_%d" % i] >= 0).sum()) for i in range(2)], [int(out["sn_%d" % i]) for i in range(2)])

out = {}
cfgs = [(895, 3000, 2000, 10), (896, 6000, 1000, 4)]
for i, (idx, nmp, nkp, th) in enumerate(cfgs):
    pre = np.where(np.random.default_rng(idx).random(nkp) < 0.1, -2, -1).astype(np.int32)
    r = R.ref_search_by_projection_sim3(fuse_frame(idx, nmp, nkp), (-13.7, -9.2, 661.3, 492.8), th, pre)
    out["cfg_%d" % i] = np.array([idx, nmp, nkp, th]); out["n_%d" % i] = r[0]; out["m_%d" % i] = r[1]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_match_projsim3.npz"), **out)
print("ref_match_projsim3.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])

out = {}
import tempfile  # noqa: E402
from weiner_slamit_v2_b200.workloads import synthetic_vocabulary, vocabulary_features, write_vocabulary_text  # noqa: E402
cfgs = [(0, 10, 4, 2000, 4), (1, 6, 5, 1500, 2), (2, 3, 6, 800, 4)]
tmp = tempfile.mkdtemp()
for i, (vi, k, L, n, lu) in enumerate(cfgs):
    voc = synthetic_vocabulary(vi, k, L)
    path = os.path.join(tmp, "voc%d.txt" % vi)
    write_vocabulary_text(path, voc)
    r = R.ref_bow_transform(path, vocabulary_features(vi, voc, n), lu)
    out["cfg_%d" % i] = np.array([vi, k, L, n, lu])
    for key in ("word", "value", "node", "start", "feat"):
        out["%s_%d" % (key, i)] = r[key]
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_bow_transform.npz"), **out)
print("ref_bow_transform.npz", [len(out["word_%d" % i]) for i in range(len(cfgs))])

# Frame::ComputeStereoMatches: the reference's own extractor (keypoints, descriptors, pyramids of both images) feeding
# the reference's own stereo matcher; the images are regenerated from their seed
from weiner_slamit_v2_b200.frames import stereo_right_frame  # noqa: E402
out = {}
cfgs = [(0, 640, 480, 1000, 0.1, 40.0), (5, 640, 480, 1000, 0.5, 20.0), (7, 752, 480, 1200, 0.11, 47.9)]
for i, (idx, w, h, nf, mb, mbf) in enumerate(cfgs):
    left = synthetic_frame(idx, w, h); right = stereo_right_frame(left, idx)
    side = []
    for img in (left, right):
        ex = R.RefExtractor(nf, 1.2, 8, 20, 7)
        k, d = ex(img)
        side.append((k, d, [ex.level_pixels(l) for l in range(8)], ex.tables()))
    (kl, dl, lp, tb), (kr, dr, rp, _) = side
    ur, dep, cnt = R.ref_compute_stereo_matches(kl, dl, kr, dr, tb["scale"], tb["inv_scale"], lp, rp, mb, mbf)
    out["cfg_%d" % i] = np.array([idx, w, h, nf, mb, mbf], np.float64)
    out["ur_%d" % i] = ur; out["depth_%d" % i] = dep; out["n_%d" % i] = cnt
    out["kl_sha_%d" % i] = digest(kl); out["kr_sha_%d" % i] = digest(kr)
out["count"] = len(cfgs)
np.savez_compressed(os.path.join(OUT, "ref_stereo_matches.npz"), **out)
print("ref_stereo_matches.npz", [int(out["n_%d" % i]) for i in range(len(cfgs))])
