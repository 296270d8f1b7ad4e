"""Small end-to-end exercise for compute-sanitizer (memcheck): a few odd-sized frames through every kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dani_slam_b200 import orbx, synth
for (w, h, nf) in [(641, 479, 600), (320, 240, 300), (752, 480, 1200)]:
    ex = orbx.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=3)
    imgs = np.stack([synth.parity_frame(s, w, h) for s in range(3)])
    n, mono, k, d = ex.extract_batch(imgs, (0, 0))
    print(w, h, n.tolist())
    ex.mvDynamicArea = [(10, 10, 100, 100)]
    print(ex(imgs[0], None, (0, 1000))[0])
    ex.close()
m = orbx.ORBmatcher(0.7, True)
q, db = synth.knn_case(300, 5001, seed=3)
idx, dist = m.knnMatch(q, db)
print(idx[:2].tolist(), m.ratio_test(dist).sum())
print("done")
