import sys, numpy as np
sys.path.insert(0, '.')
from cerebro_b200.frontend import FrontEnd
from oracle import stereo
from tests.synth_stereo import stereo_scene
fe = FrontEnd(max_pairs=1, max_features=512)
pairs = [stereo_scene(150, 260, kind, seed=20 + kind) for kind in range(2)]
L, R = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
d = fe.stereo_bm(L, R, ndisp=64, wsz=21)
ok = all(np.array_equal(d[k], stereo.stereo_bm(L[k], R[k], 64, 21)) for k in range(2))
l, r = stereo_scene(97, 211, 1, seed=3)
ok2 = np.array_equal(fe.stereo_bm(l, r, 16, 5), stereo.stereo_bm(l, r, 16, 5))
Q = np.array([[1, 0, 0, -130.5], [0, 1, 0, -75.25], [0, 0, 0, 421.3], [0, 0, 8.33, 0.0]])
p3 = fe.disparity_to_3d(d, Q)
ok3 = all(np.array_equal(p3[k], stereo.disparity_to_3d(d[k], -130.5, -75.25, 421.3, 8.33, 0.0)) for k in range(2))
big = stereo_scene(480, 640, 0, seed=1)
fe.stereo_bm(np.stack([big[0]] * 8), np.stack([big[1]] * 8))
print("STEREO", ok, ok2, ok3, "valid", float((d >= 0).mean()), "mismatch", int((d[0] != stereo.stereo_bm(L[0], R[0], 64, 21)).sum()), "ms_8x480x640", fe.last_stereo_ms())
