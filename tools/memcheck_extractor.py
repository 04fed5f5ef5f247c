import sys, numpy as np
sys.path[:0] = ['.', 'active-orb-slam2_b200']
from orbx import synth
from orbx.extractor import ORBextractor
from oracle import oracle_py as O
for kind, seed, w, h, nf in (("rect", 1, 640, 480, 1000), ("sparse", 2, 1241, 376, 2000), ("noise", 3, 333, 251, 300)):
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    img = synth.frame(kind, seed, w, h)
    k, d = ex.extract_batch([img, img])
    k2, d2 = O.Extractor(nf)(img)
    print(kind, w, h, len(k[0]), k[0].tobytes() == k2.tobytes() and np.array_equal(d[0], d2) and k[1].tobytes() == k2.tobytes())
    ex.close()
