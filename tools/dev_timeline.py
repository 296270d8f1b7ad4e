import sys, os, ctypes as C
os.environ["ORBX_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, '/root/repo')
import numpy as np
from dani_slam_b200 import orbx, synth
imgs = np.stack([synth.throughput_frame(i) for i in range(8)])
ex = orbx.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=8)
ex.L.orbx_debug_timeline.argtypes = [C.c_void_p, C.c_void_p]
for it in range(3):
    ex.extract_batch(imgs)
    out = np.zeros(32, np.int64)
    ex.L.orbx_debug_timeline(ex.h, out.ctypes.data_as(C.c_void_p))
    n = int(out[31]); t = out[:n]
    print("stamps", n, "deltas(cycles):", np.diff(t).tolist(), "total", int(t[-1]-t[0]))
