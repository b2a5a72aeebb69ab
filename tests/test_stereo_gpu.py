"""GPU parity of Frame::ComputeStereoMatches (S/Frame.cc:591-763) through the C ABI
(orbb200_compute_stereo_matches): against the CPU oracle on the same inputs and against the vectors the reference's own
extractor + stereo matcher produced (tests/golden/ref_stereo_matches.npz)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from weiner_slamit_v2_b200 import ORBextractor
from weiner_slamit_v2_b200 import frames as F
from weiner_slamit_v2_b200._lib import KP_DTYPE
from weiner_slamit_v2_b200.matcher import ORBmatcher

pytestmark = pytest.mark.gpu
SGOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_stereo_matches.npz")


def _pairs(indices, w=640, h=480):
    lefts = np.stack([F.synthetic_frame(i, w, h) for i in indices])
    rights = np.stack([F.stereo_right_frame(lefts[j], i) for j, i in enumerate(indices)])
    return lefts, rights


def _extract(images, nf=1000):
    ex = ORBextractor(nf, 1.2, 8, 20, 7, width=images.shape[2], height=images.shape[1], max_batch=len(images), device=0)
    kps, desc, counts = ex.extract_batch(images)
    sides = [(kps[i, :counts[i]].copy(), desc[i, :counts[i]].copy()) for i in range(len(images))]
    return ex, sides


def _oracle(sides_l, sides_r, exl, exr, mb, mbf):
    out = []
    for i, ((kl, dl), (kr, dr)) in enumerate(zip(sides_l, sides_r)):
        lp = [exl.get_level(i, l) for l in range(8)]; rp = [exr.get_level(i, l) for l in range(8)]
        out.append(O.compute_stereo_matches(kl, dl, kr, dr, exl.GetScaleFactors(), exl.GetInverseScaleFactors(), lp, rp, mb, mbf))
    return out


@pytest.mark.parametrize("mb,mbf", [(0.1, 40.0), (0.5, 20.0)])
def test_stereo_device_pyramids_match_oracle(mb, mbf):
    """Two extractor handles (left / right), pyramids read where the extractors left them in device memory."""
    lefts, rights = _pairs([0, 1, 2, 3])
    exl, sl = _extract(lefts); exr, sr = _extract(rights)
    m = ORBmatcher(device=0)
    res, nm = m.compute_stereo_matches_batch(sl, sr, exl.pyramid_view(), exr.pyramid_view(), exl.GetScaleFactors(),
                                             exl.GetInverseScaleFactors(), mb, mbf)
    assert m.last_launches == 2
    for i, (ur_o, dep_o, kept, skipped) in enumerate(_oracle(sl, sr, exl, exr, mb, mbf)):
        assert skipped == 0 and kept == nm[i] and kept > 50
        assert res[i][0].tobytes() == ur_o.tobytes() and res[i][1].tobytes() == dep_o.tobytes()


def test_stereo_host_pyramids_ragged_and_empty():
    """Host level images staged by the call; pairs with different keypoint counts, an empty right side, an empty pair."""
    lefts, rights = _pairs([4, 5, 6])
    exl, sl = _extract(lefts); exr, sr = _extract(rights)
    lp = [[exl.get_level(i, l) for l in range(8)] for i in range(3)]
    rp = [[exr.get_level(i, l) for l in range(8)] for i in range(3)]
    sl[1] = (sl[1][0][:300], sl[1][1][:300])
    sr[1] = (sr[1][0][100:700], sr[1][1][100:700])
    sr[2] = (sr[2][0][:0], sr[2][1][:0])
    sl.append((sl[0][0][:0], sl[0][1][:0])); sr.append(sr[0]); lp.append(lp[0]); rp.append(rp[0])
    m = ORBmatcher(device=0)
    sc, isc = exl.GetScaleFactors(), exl.GetInverseScaleFactors()
    res, nm = m.compute_stereo_matches_batch(sl, sr, lp, rp, sc, isc, 0.1, 40.0)
    for i in range(4):
        ur_o, dep_o, kept, _ = O.compute_stereo_matches(sl[i][0], sl[i][1], sr[i][0], sr[i][1], sc, isc, lp[i], rp[i], 0.1, 40.0)
        assert kept == nm[i]
        assert res[i][0].tobytes() == ur_o.tobytes() and res[i][1].tobytes() == dep_o.tobytes()
    assert nm[2] == 0 and nm[3] == 0 and nm[0] > 50


def test_stereo_arbitrary_keypoints_near_the_borders():
    """Keypoints that are not the extractor's: anywhere in the image, any octave, so patches leave the level images
    (the reference throws there; the oracle and the device both give such keypoints no depth)."""
    rng = np.random.default_rng(77)
    lefts, rights = _pairs([8])
    exl, sl = _extract(lefts); exr, sr = _extract(rights)
    lp = [[exl.get_level(0, l) for l in range(8)]]; rp = [[exr.get_level(0, l) for l in range(8)]]

    def scatter(side, n):
        k = np.zeros(n, KP_DTYPE)
        k["x"] = rng.uniform(0, 640, n).astype(np.float32); k["y"] = rng.uniform(0, 479.9, n).astype(np.float32)
        k["octave"] = rng.integers(0, 8, n)
        src = rng.integers(0, len(side[1]), n)
        return k, side[1][src]
    kl, dl = scatter(sl[0], 1500)
    kr = kl.copy(); kr["x"] = np.maximum(kl["x"] - rng.uniform(0, 30, len(kl)).astype(np.float32), 0); dr = dl.copy()
    m = ORBmatcher(device=0)
    sc, isc = exl.GetScaleFactors(), exl.GetInverseScaleFactors()
    res, nm = m.compute_stereo_matches_batch([(kl, dl)], [(kr, dr)], lp, rp, sc, isc, 0.08, 30.0)
    ur_o, dep_o, kept, skipped = O.compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp[0], rp[0], 0.08, 30.0)
    assert skipped > 0 and kept == nm[0]
    assert res[0][0].tobytes() == ur_o.tobytes() and res[0][1].tobytes() == dep_o.tobytes()


def test_stereo_reproduces_reference_golden_vectors():
    """The reference's own extractor + ComputeStereoMatches (tools/gen_golden.py) against extractor + stereo on the device."""
    g = np.load(SGOLDEN)
    for i in range(int(g["count"])):
        idx, w, h, nf = (int(v) for v in g["cfg_%d" % i][:4])
        mb, mbf = float(g["cfg_%d" % i][4]), float(g["cfg_%d" % i][5])
        lefts, rights = _pairs([idx], w, h)
        exl, sl = _extract(lefts, nf); exr, sr = _extract(rights, nf)
        m = ORBmatcher(device=0)
        res, nm = m.compute_stereo_matches_batch(sl, sr, exl.pyramid_view(), exr.pyramid_view(), exl.GetScaleFactors(),
                                                 exl.GetInverseScaleFactors(), mb, mbf)
        assert nm[0] == int(g["n_%d" % i])
        assert res[0][0].tobytes() == g["ur_%d" % i].tobytes() and res[0][1].tobytes() == g["depth_%d" % i].tobytes()


def test_stereo_pipeline_on_device_matches_oracle_chain():
    """Extraction of both images, keypoint views and the stereo matcher queued on the device without a host round
    trip (pipeline.StereoPipeline) against the oracle's extractor + stereo matcher."""
    import torch
    from weiner_slamit_v2_b200.pipeline import StereoPipeline
    lefts, rights = _pairs([10, 11, 12])
    pipe = StereoPipeline(max_pairs=3, mb=0.1, mbf=40.0, device=0)
    nm, ur, dep = pipe.run(torch.from_numpy(lefts).cuda(), torch.from_numpy(rights).cuda(), 3)
    pipe.sync()
    nm, ur, dep = nm.cpu().numpy(), ur.cpu().numpy(), dep.cpu().numpy()
    for i in range(3):
        side = []
        for img in (lefts[i], rights[i]):
            ex = O.OracleExtractor(1000, 1.2, 8, 20, 7)
            k, d = ex(img)
            side.append((k, d, [ex.level_pixels(l) for l in range(8)], ex.scale_factors, ex.inv_scale_factors))
        (kl, dl, lp, sc, isc), (kr, dr, rp, _, _) = side
        ur_o, dep_o, kept, skipped = O.compute_stereo_matches(kl, dl, kr, dr, sc, isc, lp, rp, 0.1, 40.0)
        assert kept == nm[i] and skipped == 0
        assert ur[i, :len(kl)].tobytes() == ur_o.tobytes() and dep[i, :len(kl)].tobytes() == dep_o.tobytes()
    pipe.close()
