"""Small end-to-end run for compute-sanitizer (developer tool): extractor (incl. ragged size and large-cell tile), matcher, N1, N2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from weiner_slamit_v2_b200 import ORBextractor
from weiner_slamit_v2_b200.frames import synthetic_frame
from weiner_slamit_v2_b200.matcher import Frame, MapPoints, ORBmatcher
from weiner_slamit_v2_b200.workloads import SCALE_FACTORS_8, init_pair, projection_frame, motion_frame
from weiner_slamit_v2_b200.pipeline import InitializationPipeline
for (w, h, nf) in ((640, 480, 1000), (641, 479, 500), (320, 240, 300)):
    ex = ORBextractor(nf, 1.2, 8, 20, 7, width=w, height=h, max_batch=3)
    fr = np.stack([synthetic_frame(i, w, h) for i in range(3)])
    k, d, c = ex.extract_batch(fr)
    print(w, h, c)
    ex.close()
rng = np.random.default_rng(0)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=2)
k, d, c = ex.extract_batch(rng.integers(0, 256, (2, 480, 640)).astype(np.uint8)); print("noise", c)
m = ORBmatcher(0.9, True, max_items=2, max_points=1000)
ps = [init_pair(i, 600, brute_force=(i == 1)) for i in range(2)]
print(m.search_for_initialization_batch([Frame(p[0], p[1], 640, 480) for p in ps], [Frame(p[2], p[3], 640, 480) for p in ps], [p[4] for p in ps], 100)[0])
kp, kd, mp = projection_frame(0, 800, 2500)
f = Frame(kp, kd, 1280, 720, SCALE_FACTORS_8)
print(m.SearchByProjection(f, MapPoints(mp["x"], mp["y"], mp["level"], mp["viewcos"], mp["desc"], mp["in_view"], mp["bad"], mp["xr"], mp["obs"]), 3.0))
wm = motion_frame(0, 700, 900)
print(m.search_by_projection_last_frame_batch([Frame(wm["cur"], wm["cdesc"], 640, 480, SCALE_FACTORS_8)], [wm], 15.0))
pipe = InitializationPipeline(max_pairs=2)
f1 = np.stack([synthetic_frame(9 + i) for i in range(2)])
nm, _, _ = pipe.run(torch.from_numpy(f1).cuda(), torch.from_numpy(np.roll(f1, 3, 2)).cuda(), 2); pipe.sync(); print("pipeline", nm.cpu().numpy())
print("done")
# stereo (row N4): device pyramids and host pyramids
from weiner_slamit_v2_b200.frames import stereo_right_frame
from weiner_slamit_v2_b200.pipeline import StereoPipeline
lf = np.stack([synthetic_frame(20 + i) for i in range(2)]); rf = np.stack([stereo_right_frame(lf[i], 20 + i) for i in range(2)])
sp = StereoPipeline(max_pairs=2)
nm, _, _ = sp.run(torch.from_numpy(lf).cuda(), torch.from_numpy(rf).cuda(), 2); sp.sync(); print("stereo pipeline", nm.cpu().numpy())
exl = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1); exr = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
kl, dl, cl = exl.extract_batch(lf[:1]); kr, dr, cr = exr.extract_batch(rf[:1])
lp = [[exl.get_level(0, l) for l in range(8)]]; rp = [[exr.get_level(0, l) for l in range(8)]]
res, nm = m.compute_stereo_matches_batch([(kl[0, :cl[0]], dl[0, :cl[0]])], [(kr[0, :cr[0]], dr[0, :cr[0]])], lp, rp,
                                         exl.GetScaleFactors(), exl.GetInverseScaleFactors(), 0.1, 40.0)
print("stereo host pyramids", nm)
print("done (stereo)")
