"""Minimal brute-force kNN run (config-4 shape) for timing / ncu: `python tools/prof_knn.py [ndb] [steps]`."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dani_slam_b200 import orbx  # noqa: E402

ndb = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
nq = 2000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
d_db = torch.randint(0, 256, (ndb, 32), dtype=torch.uint8, device=dev, generator=g)
d_q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=dev, generator=g)
idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dist = torch.empty((nq, 2), dtype=torch.int32, device=dev)
m = orbx.ORBmatcher(0.7, True, device=0)
st = torch.cuda.ExternalStream(m.stream(), device=dev)
torch.cuda.synchronize()
for _ in range(2):
    m.knn2_device(d_q.data_ptr(), nq, d_db.data_ptr(), ndb, 0, idx.data_ptr(), dist.data_ptr())
m.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(steps):
    m.knn2_device(d_q.data_ptr(), nq, d_db.data_ptr(), ndb, 0, idx.data_ptr(), dist.data_ptr())
e1.record(st)
e1.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{nq} x {ndb}: {ms:.3f} ms per scan, {nq * ndb / ms / 1e6:.0f} Gpairs/s, tensor-core launches {m.tc_launches()}")
