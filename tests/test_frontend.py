"""Correspondence front end: brute-force Hamming matching, GMS filtering and the 3D-2D / 3D-3D
set builders.  The oracle is pinned against the UNMODIFIED reference GMS matcher (oracle/_ref/libgms_ref.so, built from
/root/reference by oracle/Makefile) where it is present, against fixtures that matcher produced (tests/golden/gms_golden.npz)
everywhere, and against the installed OpenCV's BFMatcher; the CUDA path is then compared with the oracle, bit-exact."""
import numpy as np
import pytest

from oracle import gms
from tests import golden_io


def _synthetic_pair(rng, n1, n2, w, h, inlier_frac=0.6):
    """Keypoints + 256-bit descriptors of two 'images': a fraction of the query points reappears in the second image
    under a small affine motion with a few bits flipped; the rest are unrelated."""
    kp1 = np.stack([rng.uniform(0, w - 1e-3, n1), rng.uniform(0, h - 1e-3, n1)], 1).astype(np.float32)
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    kp2 = np.stack([rng.uniform(0, w - 1e-3, n2), rng.uniform(0, h - 1e-3, n2)], 1).astype(np.float32)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    m = int(min(n1, n2) * inlier_frac)
    src = rng.choice(n1, m, replace=False)
    dst = rng.choice(n2, m, replace=False)
    A = np.array([[0.98, 0.03], [-0.03, 0.98]])
    p = kp1[src] @ A.T + np.array([7.0, -4.0]) + rng.normal(0, 0.7, (m, 2))
    p[:, 0] = np.clip(p[:, 0], 0, w - 1e-3)
    p[:, 1] = np.clip(p[:, 1], 0, h - 1e-3)
    kp2[dst] = p.astype(np.float32)
    flip = rng.integers(0, 256, (m, 32), dtype=np.uint8) & rng.integers(0, 256, (m, 32), dtype=np.uint8) & rng.integers(0, 256, (m, 32), dtype=np.uint8)
    d2[dst] = d1[src] ^ flip  # ~12 % of the bits flipped
    return kp1, d1, kp2, d2


# ------------------------------------------------------------------ oracle pinning (CPU)
@pytest.mark.parametrize("case", ["easy", "hard"])
def test_oracle_matches_reference_fixtures(case):
    g = golden_io.load("gms_golden.npz")
    w, h = g[case + "_size"]
    idx, dist = gms.bf_match_hamming(g[case + "_d1"], g[case + "_d2"])
    assert np.array_equal(idx, g[case + "_train"]) and np.array_equal(dist, g[case + "_dist"])  # cv2.BFMatcher's answer
    mask, n = gms.gms_inlier_mask(g[case + "_kp1"], (w, h), g[case + "_kp2"], (w, h), np.arange(idx.size), idx)
    assert np.array_equal(mask, g[case + "_mask"]) and n == int(g[case + "_mask"].sum()) and n > 100  # the reference's answer


@pytest.mark.skipif(not gms.reference_available(), reason="oracle/_ref/libgms_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed,n1,n2,w,h", [(1, 1500, 1400, 640, 480), (2, 300, 900, 752, 480), (3, 40, 40, 320, 240), (4, 2500, 2500, 640, 480)])
def test_oracle_matches_compiled_reference(seed, n1, n2, w, h):
    rng = np.random.default_rng(seed)
    kp1, d1, kp2, d2 = _synthetic_pair(rng, n1, n2, w, h)
    idx, _ = gms.bf_match_hamming(d1, d2)
    q = np.arange(n1)
    for ws, wr in ((False, False), (True, False), (False, True), (True, True)):
        mr, nr = gms.gms_reference(kp1, (w, h), kp2, (w, h), q, idx, ws, wr)
        mo, no = gms.gms_inlier_mask(kp1, (w, h), kp2, (w, h), q, idx, ws, wr)
        assert nr == no and np.array_equal(mr, mo), (ws, wr, nr, no)


def test_bf_oracle_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    d1 = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (650, 32), dtype=np.uint8)
    d2[10] = d2[400] = d1[3]  # an exact tie: the first train index must win
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(d1, d2)
    idx, dist = gms.bf_match_hamming(d1, d2)
    assert np.array_equal(idx, [x.trainIdx for x in m]) and np.array_equal(dist, [int(x.distance) for x in m])
    assert idx[3] == 10 and dist[3] == 0


def test_collection_builders_gate_depth_and_truncate():
    K = np.array([[400.0, 0, 320.0], [0, 410.0, 240.0], [0, 0, 1]])
    img = np.zeros((480, 640, 3), dtype=np.float32)
    img[..., 2] = 5.0
    img[100, 200] = (1.0, 2.0, 0.05)  # too close
    img[101, 200] = (1.0, 2.0, 30.0)  # too far
    img[102, 200] = (1.5, 2.5, 0.1)  # boundary: kept (the test is z < 0.1)
    uv = np.array([[200.9, 100.9], [200.2, 101.7], [200.5, 102.99], [10.0, 20.0]])
    uvd = uv + 3.0
    a, b, X = gms.make_3d_2d_collection(K, uv, img, uvd)
    assert X.shape == (2, 3) and np.allclose(X[0], (1.5, 2.5, 0.1), atol=1e-7) and X[1, 2] == 5.0
    assert np.allclose(a[1], ((10.0 - 320) / 400, (20.0 - 240) / 410)) and np.allclose(b[1], ((13.0 - 320) / 400, (23.0 - 240) / 410))
    img_b = img.copy()
    img_b[23, 13, 2] = 26.0
    P, Q = gms.make_3d_3d_collection(uv, img, uvd, img_b)
    assert P.shape == (1, 3) and Q.shape == (1, 3)


# ------------------------------------------------------------------ CUDA path vs oracle (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("simt", ["0", "1"])
def test_device_match_gms_matches_golden_and_oracle(native_lib, cuda_device, monkeypatch, simt):
    """simt=0: tcgen05 int8 binary-GEMM matcher (default); simt=1: the popc kernel (CB_MATCH_SIMT=1).  Bit-exact both."""
    from cerebro_b200.frontend import FrontEnd

    monkeypatch.setenv("CB_MATCH_SIMT", simt)

    g = golden_io.load("gms_golden.npz")
    fe = FrontEnd(max_pairs=8, max_features=3000)
    rng = np.random.default_rng(11)
    pairs = [(g["easy_kp1"], g["easy_d1"], g["easy_kp2"], g["easy_d2"])]
    w, h = (int(v) for v in g["easy_size"])
    pairs.append(_synthetic_pair(rng, 1700, 2100, w, h))
    pairs.append(_synthetic_pair(rng, 130, 90, w, h, inlier_frac=0.9))
    pairs.append((np.zeros((0, 2), np.float32), np.zeros((0, 32), np.uint8)) + _synthetic_pair(rng, 5, 60, w, h)[2:])  # no query features
    pairs.append(_synthetic_pair(rng, 2999, 3000, w, h, inlier_frac=0.3))
    res = fe.match_gms([p[0] for p in pairs], [p[1] for p in pairs], [p[2] for p in pairs], [p[3] for p in pairs], (w, h), (w, h))
    assert np.array_equal(res[0]["train_idx"], g["easy_train"]) and np.array_equal(res[0]["distance"], g["easy_dist"])
    assert np.array_equal(res[0]["inliers"], g["easy_mask"]), "GMS mask differs from the reference matcher's"
    for p, r in zip(pairs, res):
        if len(p[0]) == 0:
            assert r["train_idx"].size == 0 and r["n_inliers"] == 0
            continue
        idx, dist = gms.bf_match_hamming(p[1], p[3])
        assert np.array_equal(r["train_idx"], idx) and np.array_equal(r["distance"], dist)
        mask, n = gms.gms_inlier_mask(p[0], (w, h), p[2], (w, h), np.arange(idx.size), idx)
        assert np.array_equal(r["inliers"], mask) and r["n_inliers"] == n
    assert res[1]["n_inliers"] > 300
    assert fe.last_match_ms() > 0
    # the second golden case has another image size
    w2, h2 = (int(v) for v in g["hard_size"])
    r = fe.match_gms([g["hard_kp1"]], [g["hard_d1"]], [g["hard_kp2"]], [g["hard_d2"]], (w2, h2), (w2, h2))[0]
    assert np.array_equal(r["train_idx"], g["hard_train"]) and np.array_equal(r["inliers"], g["hard_mask"])
    fe.close()


@pytest.mark.gpu
def test_device_collections_match_oracle(native_lib, cuda_device):
    from cerebro_b200.frontend import FrontEnd

    rng = np.random.default_rng(3)
    w, h = 640, 480
    K = np.array([[385.2, 0, 321.7], [0, 386.1, 238.4], [0, 0, 1]])
    pairs = [_synthetic_pair(rng, 1200, 1300, w, h), _synthetic_pair(rng, 400, 380, w, h, inlier_frac=0.8)]
    fe = FrontEnd(max_pairs=2, max_features=1500)
    res = fe.match_gms([p[0] for p in pairs], [p[1] for p in pairs], [p[2] for p in pairs], [p[3] for p in pairs], (w, h), (w, h))
    img_a = rng.uniform(-3, 3, (2, h, w, 3)).astype(np.float32)
    img_b = rng.uniform(-3, 3, (2, h, w, 3)).astype(np.float32)
    img_a[..., 2] = rng.uniform(-1, 30, (2, h, w))  # ~17 % outside the 0.1..25 m gate
    img_b[..., 2] = rng.uniform(-1, 30, (2, h, w))
    s2 = fe.make_3d_2d_collection(K, img_a)
    s3 = fe.make_3d_3d_collection(img_a, img_b)
    for p in range(2):
        u, ud = fe.matched_points(p)
        assert u.shape[0] == 3 and u.shape[1] == res[p]["n_inliers"] and np.all(u[2] == 1.0)
        a, b, X = gms.make_3d_2d_collection(K, u[:2].T, img_a[p], ud[:2].T)
        assert X.shape[0] > 50 and X.shape[0] < u.shape[1]
        assert s2[p][2].shape == X.shape and np.array_equal(s2[p][2], X)  # fp32 -> fp64 widening is exact
        assert np.allclose(s2[p][0], a, rtol=0, atol=1e-12) and np.allclose(s2[p][1], b, rtol=0, atol=1e-12)
        P, Q = gms.make_3d_3d_collection(u[:2].T, img_a[p], ud[:2].T, img_b[p])
        assert np.array_equal(s3[p][0], P) and np.array_equal(s3[p][1], Q)
    fe.close()


@pytest.mark.gpu
def test_reference_call_shape_images_in(native_lib, cuda_device):
    """StaticPointFeatureMatching::gms_point_feature_matches(imleft, imright, u, ud) with images in (ORB on the host)."""
    cv2 = pytest.importorskip("cv2")
    from cerebro_b200.frontend import StaticPointFeatureMatching

    rng = np.random.default_rng(8)
    a = cv2.GaussianBlur((rng.random((480, 640)) * 255).astype(np.uint8), (0, 0), 2.0)
    a = cv2.normalize(a, None, 0, 255, cv2.NORM_MINMAX)
    H = np.array([[1.01, 0.02, 5], [-0.02, 1.0, 3], [0, 0, 1.0]])
    b = cv2.warpPerspective(a, H, (640, 480))
    u, ud = StaticPointFeatureMatching.gms_point_feature_matches(a, b, n_orb_feat=2000)
    assert u.shape[0] == 3 and u.shape == ud.shape and u.shape[1] > 300
    proj = H @ u
    proj /= proj[2]
    err = np.linalg.norm(proj[:2] - ud[:2], axis=0)
    assert np.median(err) < 2.0  # GMS keeps the matches that follow the homography


@pytest.mark.gpu
def test_candidate_to_pose_chain(native_lib, cuda_device):
    """Row a10 of the scope table end to end on the device, minus ORB / StereoBM: features of a synthetic scene seen from
    two poses -> Hamming matches -> GMS -> 3D-2D set (Option A) -> DLS-PnP RANSAC recovers b_T_a within north_star's
    1e-3 rad / 1e-2 m, and Option C (3D-3D Umeyama RANSAC) agrees with it (ProcessedLoopCandidate.cpp:40-125 thresholds)."""
    from cerebro_b200.frontend import FrontEnd
    from cerebro_b200.pnp import PnpBatch, default_params
    from oracle import dls_pnp

    rng = np.random.default_rng(12)
    w, h = 640, 480
    K = np.array([[420.0, 0, 320.0], [0, 420.0, 240.0], [0, 0, 1]])
    # frame a: a smooth depth map, 3-D image = back-projected pixels (what the stereo pair's reprojectImageTo3D gives)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    depth = 6.0 + 2.0 * np.sin(xs / 90.0) + 1.5 * np.cos(ys / 70.0)
    img_a = np.stack([(xs - K[0, 2]) / K[0, 0] * depth, (ys - K[1, 2]) / K[1, 1] * depth, depth], -1).astype(np.float32)
    # true relative pose b_T_a: ~4 degrees, 25 cm
    ang = np.deg2rad([2.5, -3.0, 1.5])
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    R, t = Rz @ Ry @ Rx, np.array([0.2, -0.1, 0.12])
    n = 2500
    kp1 = np.stack([rng.uniform(20, w - 20, n), rng.uniform(20, h - 20, n)], 1).astype(np.float32)
    # integer pixel positions so that the (int)-truncated depth lookup is the feature's own depth
    kp1 = np.floor(kp1).astype(np.float32)
    Xa = img_a[kp1[:, 1].astype(int), kp1[:, 0].astype(int)].astype(np.float64)
    Xb = Xa @ R.T + t
    proj = Xb @ K.T
    kp2_true = proj[:, :2] / proj[:, 2:3]
    inside = (kp2_true[:, 0] > 1) & (kp2_true[:, 0] < w - 2) & (kp2_true[:, 1] > 1) & (kp2_true[:, 1] < h - 2)
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    # frame b sees 80 % of the inside points (descriptors with ~10 % flipped bits) plus unrelated features
    seen = np.nonzero(inside)[0]
    seen = seen[rng.random(seen.size) < 0.8]
    flip = rng.integers(0, 256, (seen.size, 32), dtype=np.uint8) & rng.integers(0, 256, (seen.size, 32), dtype=np.uint8) & rng.integers(0, 256, (seen.size, 32), dtype=np.uint8)
    kp2 = np.concatenate([kp2_true[seen] + rng.normal(0, 0.15, (seen.size, 2)), np.stack([rng.uniform(0, w - 1, 700), rng.uniform(0, h - 1, 700)], 1)]).astype(np.float32)
    d2 = np.concatenate([d1[seen] ^ flip, rng.integers(0, 256, (700, 32), dtype=np.uint8)])
    perm = rng.permutation(kp2.shape[0])
    kp2, d2 = kp2[perm], d2[perm]
    # frame b's own 3-D image (for Option C): depth of the scene seen from b is not analytic; give b a 3-D image that is
    # exact at the feature positions
    img_b = np.zeros((h, w, 3), dtype=np.float32)
    img_b[..., 2] = 50.0  # outside the gate everywhere else
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    pos = kp2[inv[: seen.size]]
    img_b[pos[:, 1].astype(int), pos[:, 0].astype(int)] = Xb[seen].astype(np.float32)

    fe = FrontEnd(max_pairs=1, max_features=4000)
    r = fe.match_gms([kp1], [d1], [kp2], [d2], (w, h), (w, h))[0]
    assert r["n_inliers"] > 800  # pf_matches > 800 is what makeLoopEdgeMsgWithConsistencyCheck asks for
    uv_a, uv_b, X = fe.make_3d_2d_collection(K, img_a[None])[0]
    assert X.shape[0] > 800
    pb = PnpBatch(max_candidates=2, max_points_total=10000, max_hypotheses=50)
    out = pb.solve([X], [uv_b], default_params(seed=5))
    T = out["T"][0]
    Ttrue = np.eye(4)
    Ttrue[:3, :3], Ttrue[:3, 3] = R, t
    e_rot, e_t = dls_pnp.pose_error(T, Ttrue)
    assert out["confidence"][0] > 0.9 and e_rot < 1e-3 and e_t < 1e-2, (e_rot, e_t, out["confidence"][0])
    P, Q = fe.make_3d_3d_collection(img_a[None], img_b[None])[0]
    assert P.shape[0] > 500
    icp = pb.icp([P], [Q], default_params(seed=6, error_thresh=0.1))
    e_rot_c, e_t_c = dls_pnp.pose_error(icp["T"][0], Ttrue)
    assert e_rot_c < 5e-3 and e_t_c < 3e-2, (e_rot_c, e_t_c)
    # Option B's set (frames swapped) against the oracle, then the whole a10 orchestration -> a published LoopEdge
    u, ud = fe.matched_points(0)
    ud_n, u_n, Xb_o = gms.make_3d_2d_collection(K, ud[:2].T, img_b, u[:2].T)
    uv_a2, uv_b2, Xb_d = fe.make_3d_2d_collection(K, img_b[None], swapped=True)[0]
    assert Xb_d.shape == Xb_o.shape and Xb_d.shape[0] > 500 and np.array_equal(Xb_d, Xb_o)
    assert np.allclose(uv_a2, u_n, rtol=0, atol=1e-12) and np.allclose(uv_b2, ud_n, rtol=0, atol=1e-12)
    from cerebro_b200.loop_detector import consistent_pose_compute

    cand, edge = consistent_pose_compute(fe, pb, K, img_a[None], img_b[None], [r], [(500.0, 100.0)], [(812, 140)], seed=3)[0]
    assert cand is not None and cand.pf_matches == r["n_inliers"] and len(cand.opX_b_T_a) == 3
    for T_op in cand.opX_b_T_a:
        er, et = dls_pnp.pose_error(T_op, Ttrue)
        assert er < 5e-3 and et < 3e-2, (er, et)
    assert edge is not None and edge.timestamp0 == 500.0 and edge.timestamp1 == 100.0
    assert edge.weight == pytest.approx(max(cand.opX_goodness)) and edge.description.startswith("812<=>140")
    er, et = dls_pnp.pose_error(edge.pose_1T0, Ttrue)
    assert er < 1e-3 and et < 1e-2
    # same candidate seen 5 s apart: the consistency check refuses it (ProcessedLoopCandidate.cpp:49-56)
    _, edge2 = consistent_pose_compute(fe, pb, K, img_a[None], img_b[None], [r], [(105.0, 100.0)], seed=3)[0]
    assert edge2 is None
    # every depth outside the 0.1 .. 25 m gate: the solvers refuse (confidence -1, identity pose on the device); the
    # reference's uninitialised matrices fail its checks -- here the candidate must be rejected, not published as identity
    far_a, far_b = img_a.copy(), img_b.copy()
    far_a[..., 2] = 60.0
    far_b[..., 2] = 60.0
    cand3, edge3 = consistent_pose_compute(fe, pb, K, far_a[None], far_b[None], [r], [(500.0, 100.0)], seed=3)[0]
    assert cand3 is None and edge3 is None
    fe.close()
    pb.close()


@pytest.mark.gpu
def test_raw_stereo_pairs_to_loop_edge_all_on_device(native_lib, cuda_device):
    """Scope row a10 end to end (Cerebro.cpp:1414-1771): two keyframes' raw stereo pairs -> remap -> StereoBM -> 3-D image, ORB
    -> matcher -> GMS -> the three sets -> Options A / B / C -> consistency check -> LoopEdge, every stage behind the C ABI.
    A slanted textured plane seen by a rectified pair (f 400 px, baseline 0.12 m, depth 3 - 4.8 m); keyframe b is the same
    rig moved 0.06 m to the right.  The same chain with OpenCV's ORB and StereoBM spliced in must give the same edge."""
    from cerebro_b200.features import Features
    from cerebro_b200.frontend import FrontEnd
    from cerebro_b200.loop_detector import consistent_pose_compute, process_loop_candidates_from_raw_stereo
    from cerebro_b200.pnp import PnpBatch
    from tests.synth_stereo import _blur

    h, w, f, B = 480, 640, 400.0, 0.12
    rng = np.random.default_rng(7)
    pad = 80
    tex = _blur(rng.random((h, w + 2 * pad)), 1.0)
    tex = (tex - tex.min()) / (tex.max() - tex.min()) * 255.0
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    d = 10.0 + 6.0 * xs / w  # disparity of the plane, defined on the target pixel grid

    def view(shift):  # image whose pixel x shows texture position x + shift(x)
        src = xs + pad + shift
        x0 = np.floor(src).astype(int)
        a = src - x0
        rows = np.arange(h)[:, None]
        img = (1 - a) * tex[rows, np.clip(x0, 0, tex.shape[1] - 1)] + a * tex[rows, np.clip(x0 + 1, 0, tex.shape[1] - 1)]
        return np.clip(np.rint(img + rng.normal(0, 1.0, img.shape)), 0, 255).astype(np.uint8)

    left_a, right_a = view(0 * d), view(d)            # right(x) = left(x + d)
    left_b, right_b = view(0.5 * d), view(1.5 * d)    # the rig moved half a baseline to the right
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1.0]])
    Q = np.array([[1, 0, 0, -w / 2], [0, 1, 0, -h / 2], [0, 0, 0, f], [0, 0, 1.0 / B, 0]])  # cv::stereoRectify's Q, Tx = -B
    # identity warps in the two remap slots exercise the rectification stage (bit-exact pass-through)
    ident_x, ident_y = np.tile(np.arange(w, dtype=np.float32), (h, 1)), np.tile(np.arange(h, dtype=np.float32)[:, None], (1, w))
    feat = Features(h, w, max_images=2, max_keypoints=8192)
    feat.set_remap(0, ident_x, ident_y)
    feat.set_remap(1, ident_x, ident_y)
    fe = FrontEnd(max_pairs=1, max_features=8192)
    pb = PnpBatch(max_candidates=2, max_points_total=20000, max_hypotheses=50)
    out, matches = process_loop_candidates_from_raw_stereo(feat, fe, pb, K, Q, left_a[None], right_a[None], left_b[None], right_b[None],
                                                           [(500.0, 100.0)], [(40, 7)], warp_slots=((0, 1), (0, 1)), seed=3)
    cand, edge = out[0]
    assert matches[0]["n_inliers"] > 800, matches[0]["n_inliers"]
    assert cand is not None and edge is not None
    T = edge.pose_1T0
    assert np.abs(T[:3, :3] - np.eye(3)).max() < 0.02 and np.abs(T[:3, 3] - np.array([-0.06, 0, 0])).max() < 0.03, T
    assert edge.timestamp0 == 500.0 and edge.timestamp1 == 100.0 and edge.description.startswith("40<=>7")
    try:
        import cv2
    except ImportError:
        return
    # the same chain with OpenCV's own ORB and StereoBM outputs spliced in
    orb = cv2.ORB_create(5000)
    orb.setFastThreshold(0)
    k1, d1 = orb.detectAndCompute(left_a, None)
    k2, d2 = orb.detectAndCompute(left_b, None)
    bm = cv2.StereoBM_create(64, 21)
    img3d = fe.disparity_to_3d(np.stack([bm.compute(left_a, right_a), bm.compute(left_b, right_b)]), Q)
    m2 = fe.match_gms([np.array([k.pt for k in k1], np.float32)], [d1], [np.array([k.pt for k in k2], np.float32)], [d2], (w, h), (w, h))
    cand2, edge2 = consistent_pose_compute(fe, pb, K, img3d[:1], img3d[1:], m2, [(500.0, 100.0)], [(40, 7)], seed=3)[0]
    assert m2[0]["n_inliers"] == matches[0]["n_inliers"]
    assert edge2 is not None and np.array_equal(edge2.pose_1T0, edge.pose_1T0) and edge2.weight == edge.weight
    feat.close()
    fe.close()
    pb.close()
