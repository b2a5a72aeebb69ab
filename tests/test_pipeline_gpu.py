"""Frame glue (N1) and the device-resident extractor -> frame -> matcher pipeline against the oracle chain."""
import numpy as np
import pytest

import oracle_lib as O
from weiner_slamit_v2_b200.frames import synthetic_frame
from weiner_slamit_v2_b200.matcher import ORBmatcher
from weiner_slamit_v2_b200.pipeline import REFERENCE_DIST, REFERENCE_K, InitializationPipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dist", [REFERENCE_DIST, [0.262383, -0.953104, -0.005358, 0.002628, 0.0], [-0.28, 0.07, 2e-4, 2e-5, 0.0]])
def test_undistort_points_and_bounds_match_oracle(dist):
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-50, 700, 30000), rng.uniform(-50, 530, 30000)], 1).astype(np.float32)
    m = ORBmatcher()
    got = m.undistort_points(pts, REFERENCE_K, dist)
    assert np.array_equal(got, O.undistort_points(pts, REFERENCE_K, dist))
    assert np.array_equal(m.image_bounds(640, 480, REFERENCE_K, dist), O.image_bounds(640, 480, REFERENCE_K, dist))
    zero = [0, 0, 0, 0, 0]
    assert np.array_equal(m.image_bounds(640, 480, REFERENCE_K, zero), np.array([0, 0, 640, 480], np.float32))


@pytest.mark.parametrize("dist", [REFERENCE_DIST, np.zeros(5, np.float32)])
def test_device_pipeline_equals_oracle_chain(dist):
    import torch
    pairs = 3
    f1 = np.stack([synthetic_frame(50 + i) for i in range(pairs)])
    f2 = np.stack([np.roll(f, (2, 4), (0, 1)) for f in f1])                  # 4 px right, 2 px down
    pipe = InitializationPipeline(max_pairs=pairs, dist=dist)
    nm, m12, prev = pipe.run(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), pairs)
    pipe.sync()
    nm, m12, prev = nm.cpu().numpy(), m12.cpu().numpy(), prev.cpu().numpy()
    orc = O.OracleExtractor()
    bounds = O.image_bounds(640, 480, REFERENCE_K, dist)
    assert np.array_equal(bounds, pipe.bounds)
    total = 0
    for i in range(pairs):
        k1, d1 = orc(f1[i]); k2, d2 = orc(f2[i])
        if dist[0] != 0:
            for k in (k1, k2):
                u = O.undistort_points(np.stack([k["x"], k["y"]], 1), REFERENCE_K, dist)
                k["x"], k["y"] = u[:, 0], u[:, 1]
        p0 = np.stack([k1["x"], k1["y"]], 1)
        on, om12, opm = O.search_for_initialization(k1, d1, k2, d2, p0, bounds, 0.9, True, 100)
        n1 = len(k1)
        assert nm[i] == on and np.array_equal(m12[i, :n1], om12) and np.array_equal(prev[i, :n1], opm), i
        total += on
    assert total > 50
    pipe.close()
