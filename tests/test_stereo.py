"""Stereo depth step: the oracle is pinned against the installed OpenCV's StereoBM (the reference's
own dependency, src/utils/CameraGeometry.cpp:81) and the fixtures it produced."""
import numpy as np
import pytest

from oracle import stereo
from tests import golden_io
from tests.synth_stereo import stereo_scene


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_stereo_oracle_matches_opencv_fixtures(name):
    g = golden_io.load("stereo_golden.npz")
    h, w, nd, ws, kind, seed = (int(v) for v in g[name + "_cfg"])
    left, right = stereo_scene(h, w, kind, seed)
    d = stereo.stereo_bm(left, right, ndisp=nd, wsz=ws)
    assert d.dtype == np.int16 and np.array_equal(d, g[name + "_disp"])
    assert (d >= 0).mean() > 0.2


@pytest.mark.parametrize("h,w,nd,ws", [(240, 320, 64, 21), (121, 203, 64, 21), (97, 211, 16, 5), (200, 330, 128, 15)])
def test_stereo_oracle_matches_live_opencv(h, w, nd, ws):
    cv2 = pytest.importorskip("cv2")
    for kind in range(3):
        left, right = stereo_scene(h, w, kind, seed=10 + kind)
        assert np.array_equal(stereo.stereo_bm(left, right, ndisp=nd, wsz=ws), cv2.StereoBM_create(nd, ws).compute(left, right)), (kind,)


def test_disparity_to_3d_restates_the_reference_loop():
    """CameraGeometry.cpp:500-520 evaluated element by element in the same precisions."""
    rng = np.random.default_rng(0)
    disp = rng.integers(-16, 1024, (7, 9)).astype(np.int16)
    Q03, Q13, Q23, Q32, Q33 = -318.2, -241.7, 421.3, 8.33, 0.0
    out = stereo.disparity_to_3d(disp, Q03, Q13, Q23, Q32, Q33)
    for i in range(7):
        for j in range(9):
            pw = np.float32(1.0 / (float(np.float32(disp[i, j])) / 16.0 * float(np.float32(Q32)) + float(np.float32(Q33)) + 1e-6))
            assert out[i, j, 0] == (np.float32(j) + np.float32(Q03)) * pw
            assert out[i, j, 1] == (np.float32(i) + np.float32(Q13)) * pw
            assert out[i, j, 2] == np.float32(Q23) * pw
    # a metric sanity check: disparity 16 px at f = 420, baseline 0.12 m -> depth f b / d
    f, b = 420.0, 0.12
    z = stereo.disparity_to_3d(np.full((1, 1), 16 * 16, np.int16), 0, 0, f, 1.0 / b, 0.0)[0, 0, 2]
    assert abs(z - f * b / 16.0) < 1e-3
