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


def _emul():
    import ctypes

    from cerebro_b200 import build

    build.build()
    lib = ctypes.CDLL(build.STEREO_EMUL)
    P = ctypes.c_void_p
    lib.sbm_emulate.argtypes = [P, P] + [ctypes.c_int] * 6 + [P]
    lib.sbm_emulate_3d.argtypes = [P, ctypes.c_int, ctypes.c_int] + [ctypes.c_float] * 5 + [P]
    return lib


@pytest.mark.parametrize("h,w,nd,ws,seg,stripe", [(120, 200, 32, 9, 32, 60), (97, 211, 64, 21, 32, 60), (150, 260, 64, 21, 17, 33),
                                                  (121, 203, 16, 5, 64, 500), (60, 90, 64, 21, 32, 60)])
def test_kernel_bodies_emulated_on_cpu_match_oracle(h, w, nd, ws, seg, stripe):
    """csrc/stereo_core.h (the per-thread bodies the CUDA kernels run) walked over the kernels' (block, thread) space by
    host/stereo_emul.cpp -- (32, 60) is the kernels' own segment / stripe size -- against the oracle, bit-exact."""
    lib = _emul()
    for kind in range(3):
        left, right = stereo_scene(h, w, kind, seed=5 + kind)
        out = np.zeros((h, w), np.int16)
        lib.sbm_emulate(left.ctypes.data, right.ctypes.data, h, w, nd, ws, seg, stripe, out.ctypes.data)
        assert np.array_equal(out, stereo.stereo_bm(left, right, ndisp=nd, wsz=ws)), kind
    disp = stereo.stereo_bm(*stereo_scene(h, w, 0, seed=5), ndisp=nd, wsz=ws)
    o3 = np.zeros((h, w, 3), np.float32)
    lib.sbm_emulate_3d(disp.ctypes.data, h, w, -160.5, -120.25, 421.3, 8.33, 0.0, o3.ctypes.data)
    assert np.array_equal(o3, stereo.disparity_to_3d(disp, -160.5, -120.25, 421.3, 8.33, 0.0))


@pytest.mark.gpu
def test_device_stereo_matches_oracle(native_lib, cuda_device):
    """cb_frontend_stereo_bm / cb_frontend_disparity_to_3d against the oracle (itself bit-exact against cv2.StereoBM)."""
    from cerebro_b200.frontend import FrontEnd

    fe = FrontEnd(max_pairs=1, max_features=512)
    pairs = [stereo_scene(150, 260, kind, seed=20 + kind) for kind in range(3)]
    L, R = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    d = fe.stereo_bm(L, R, ndisp=64, wsz=21)
    for k in range(3):
        assert np.array_equal(d[k], stereo.stereo_bm(L[k], R[k], ndisp=64, wsz=21)), k
    assert fe.last_stereo_ms() > 0 and (d >= 0).mean() > 0.3
    left, right = stereo_scene(97, 211, 1, seed=3)
    d2 = fe.stereo_bm(left, right, ndisp=16, wsz=5)
    assert np.array_equal(d2, stereo.stereo_bm(left, right, ndisp=16, wsz=5))
    g = golden_io.load("stereo_golden.npz")  # cv2's own answer
    h, w, nd, ws, kind, seed = (int(v) for v in g["b_cfg"])
    assert np.array_equal(fe.stereo_bm(*stereo_scene(h, w, kind, seed), ndisp=nd, wsz=ws), g["b_disp"])
    Q = np.array([[1, 0, 0, -130.5], [0, 1, 0, -75.25], [0, 0, 0, 421.3], [0, 0, 8.33, 0.0]])
    p3 = fe.disparity_to_3d(d, Q)
    for k in range(3):
        assert np.array_equal(p3[k], stereo.disparity_to_3d(d[k], Q[0, 3], Q[1, 3], Q[2, 3], Q[3, 2], Q[3, 3])), k
    with pytest.raises(Exception):
        fe.stereo_bm(left, right, ndisp=20, wsz=5)
    fe.close()


@pytest.mark.gpu
def test_device_stereo_full_size_and_both_vertical_passes(native_lib, cuda_device, monkeypatch):
    """480 x 640 (the bench / EuRoC-sized frames): the warp-per-column vertical pass (default), the block-per-column one
    (CB_SBM_V1=1) and the installed OpenCV -- or, without cv2, the oracle on a crop-free 480 x 640 pair -- give the same
    int16 image; textureless band, disparity jumps and the ROI border included."""
    from cerebro_b200.frontend import FrontEnd

    pairs = [stereo_scene(480, 640, kind, seed=40 + kind) for kind in range(3)]
    L, R = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
    out = []
    for v1 in ("0", "1"):
        monkeypatch.setenv("CB_SBM_V1", v1)
        fe = FrontEnd(max_pairs=1, max_features=512)
        out.append(fe.stereo_bm(L, R, ndisp=64, wsz=21))
        fe.close()
    assert np.array_equal(out[0], out[1])
    assert (out[0] >= 0).mean() > 0.3 and (out[0] == -16).mean() > 0.05
    try:
        import cv2

        bm = cv2.StereoBM_create(64, 21)
        ref = np.stack([bm.compute(L[k], R[k]) for k in range(3)])
    except ImportError:
        ref = np.stack([stereo.stereo_bm(L[k], R[k], ndisp=64, wsz=21) for k in range(3)])
    assert np.array_equal(out[0], ref)
