"""Minimal device-resident extractor run for ncu captures: `python tools/prof_extract.py [batch] [steps] [workload]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from dani_slam_b200 import orbx  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bench.set_workload(sys.argv[3] if len(sys.argv) > 3 else "tum1")
W, H, NF = bench.W_IMG, bench.H_IMG, bench.NFEAT
cap = NF + 2 * bench.LEVELS + 8
dev = torch.device("cuda", 0)
d_frames = torch.from_numpy(bench.make_frames(B, 0)).to(dev)
d_kps = torch.zeros((B, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(B, dtype=torch.int32, device=dev)
d_mono = torch.zeros(B, dtype=torch.int32, device=dev)
ex = orbx.ORBextractor(NF, bench.SCALE, bench.LEVELS, bench.INI_TH, bench.MIN_TH, device=0, max_width=W, max_height=H, max_batch=B)
ex.set_profiling(True)       # everything on the main stream, one launch per stage
for _ in range(steps):
    ex.extract_batch_device(d_frames.data_ptr(), H * W, B, H, W, W, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr(), d_mono.data_ptr())
ex.sync()
ms, calls = ex.stage_ms()
print({k: round(v / max(calls, 1), 4) for k, v in ms.items()}, "keypoints", int(d_n.min()), int(d_n.max()))
