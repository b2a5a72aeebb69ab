"""Per-stage device times of the extractor (developer tool): 20 device-resident calls of batch B after warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from weiner_slamit_v2_b200 import ORBextractor
from weiner_slamit_v2_b200.frames import synthetic_frame
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = np.stack([synthetic_frame(i) for i in range(16)])
frames = np.concatenate([base] * (B // 16 + 1))[:B]
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
d = torch.from_numpy(frames).cuda()
ex.set_profiling(True)
for _ in range(5):
    ex.extract_device(d, B, 640, 640 * 480); ex.sync()
acc = np.zeros(5); N = 20
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for _ in range(N):
    ex.extract_device(d, B, 640, 640 * 480); acc += ex.stage_ms()
ex.set_profiling(False)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for _ in range(N):
    ex.extract_device(d, B, 640, 640 * 480)
ex.sync()
dt = (time.perf_counter() - t) / N * 1e3
print("stages ms [pyramid fast quadtree blur describe]:", np.round(acc / N, 4), "sum", round(float(acc.sum() / N), 4), "| step", round(dt, 4), "ms ->", round(B / dt * 1e3), "frames/s")
