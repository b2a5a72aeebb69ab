"""First-contact GPU script: parity summary + rough per-stage timing (developer tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import oracle_lib as O
from weiner_slamit_v2_b200 import ORBextractor
from weiner_slamit_v2_b200.frames import synthetic_frame

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
frames = np.stack([synthetic_frame(i) for i in range(min(B, 8))])
frames = np.concatenate([frames] * ((B + len(frames) - 1) // len(frames)))[:B]
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
orc = O.OracleExtractor(1000, 1.2, 8, 20, 7)
t = time.time(); kps, desc, counts = ex.extract_batch(frames); print("host call %.1f ms" % ((time.time() - t) * 1e3), counts[:8])
for f in range(min(B, 4)):
    ko, do = orc(frames[f])
    n = counts[f]
    msg = []
    for l in range(8):
        if not np.array_equal(ex.get_level(f, l), orc.level_pixels(l)): msg.append("pyr%d" % l)
        cx, cy, cs = ex.get_candidates(f, l); oc = orc.level_candidates(l)
        if not (len(cx) == len(oc) and np.array_equal(cx, oc["x"].astype(np.int32)) and np.array_equal(cy, oc["y"].astype(np.int32)) and np.array_equal(cs, oc["response"].astype(np.int32))): msg.append("cand%d(%d vs %d)" % (l, len(cx), len(oc)))
        kx, ky, ks = ex.get_level_keypoints(f, l); lk = orc.level_keypoints(l)
        if not (len(kx) == len(lk) and np.array_equal(kx, lk["x"].astype(np.int32)) and np.array_equal(ky, lk["y"].astype(np.int32))): msg.append("qt%d(%d vs %d)" % (l, len(kx), len(lk)))
        ob = orc.level_blurred(l)
        if ob is not None and not np.array_equal(ex.get_level(f, l, True), ob): msg.append("blur%d" % l)
    same = n == len(ko) and kps[f, :n].tobytes() == ko.tobytes() and np.array_equal(desc[f, :n], do)
    if not same and n == len(ko):
        for name in ko.dtype.names:
            d = int((kps[f, :n][name] != ko[name]).sum())
            if d: msg.append("%s:%d" % (name, d))
        msg.append("descrows:%d" % int((desc[f, :n] != do).any(1).sum()))
    print("frame", f, "n", n, len(ko), "IDENTICAL" if same else "DIFF", msg)
# device-resident timing
d = torch.from_numpy(frames).cuda()
st = torch.cuda.ExternalStream(ex.stream)
ex.set_profiling(True)
for it in range(3):
    ex.extract_device(d, B, 640, 640 * 480); ex.sync()
acc = np.zeros(5)
for it in range(5):
    ex.extract_device(d, B, 640, 640 * 480); acc += ex.stage_ms()
print("stage ms (pyramid, fast, quadtree, blur, describe):", np.round(acc / 5, 4))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record()
    for it in range(10):
        ex.extract_device(d, B, 640, 640 * 480)
    e1.record()
ex.sync(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("device batch %d: %.3f ms/batch -> %.0f frames/s, launches %d" % (B, ms, B / ms * 1e3, ex.last_launches))
