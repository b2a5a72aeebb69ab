"""Minimal profiling target: 3 device-resident extraction calls of batch 256 (developer tool for ncu)."""
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
for _ in range(3):
    ex.extract_device(d, B, 640, 640 * 480); ex.sync()
print("ok")
