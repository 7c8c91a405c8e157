"""Option C of the verifier (3D-3D Umeyama RANSAC) on the GPU vs the oracle, and the three-way consistency
check that turns the three poses into a LoopEdge (host logic, CPU test)."""
import numpy as np
import pytest

from oracle import dls_pnp as D
from oracle import umeyama as U


@pytest.mark.gpu
def test_icp_ransac_matches_oracle(native_lib, cuda_device):
    from cerebro_b200.pnp import PnpBatch, StaticTheiaPoseCompute, default_params

    rng = np.random.default_rng(12)
    cands = [U.synth_3d3d(rng, n=n, outlier_frac=o) for n, o in ((200, 0.2), (64, 0.0), (333, 0.3), (21, 0.1), (19, 0.0))]
    pb = PnpBatch(max_candidates=8, max_points_total=2000, max_hypotheses=64)
    for adaptive in (1, 0):
        prm = default_params(error_thresh=0.1, seed=77, max_iterations=50, adaptive=adaptive)
        r = pb.icp([c[0] for c in cands], [c[1] for c in cands], prm)
        for ci, (a, b, T) in enumerate(cands):
            o = U.ransac_icp(a, b, U.sample_table(77, ci, 50, len(a)) if len(a) >= 20 else np.zeros((50, 10), int),
                             D.RansacParameters(error_thresh=0.1, sample_size=10, adaptive=bool(adaptive)))
            assert abs(float(r["confidence"][ci]) - o["confidence"]) < 1e-6, (adaptive, ci)
            assert r["num_iterations"][ci] == o["num_iterations"] and r["n_inliers"][ci] == o["n_inliers"]
            assert r["best_hyp"][ci] == o["best_hyp"]
            e = D.pose_error(r["T"][ci], o["T"])
            assert e[0] < 1e-3 and e[1] < 1e-2
    # explicit sample table + reference call shape
    a, b, T = cands[0]
    tab = U.sample_table(5, 0, 50, 200)
    r = pb.icp([a], [b], default_params(error_thresh=0.1), samples=tab[None])
    o = U.ransac_icp(a, b, tab)
    assert r["best_hyp"][0] == o["best_hyp"]
    uvd_T_uv = np.eye(4)
    msg = []
    conf = StaticTheiaPoseCompute.P3P_ICP(a, b, uvd_T_uv, msg)
    assert conf > 0.5 and msg
    e = D.pose_error(uvd_T_uv, T)
    assert e[0] < 5e-3 and e[1] < 5e-2
    assert StaticTheiaPoseCompute.P3P_ICP(a[:10], b[:10], uvd_T_uv) == -1.0
    pb.close()


def test_consistency_check_matches_oracle():
    from cerebro_b200.loop_detector import ProcessedLoopCandidate, R2ypr

    rng = np.random.default_rng(3)
    n_pub = 0
    for trial in range(300):
        T = np.eye(4)
        T[:3, :3] = D.ypr_to_R(*np.deg2rad(rng.uniform(-40, 40, 3)))
        T[:3, 3] = rng.uniform(-2, 2, 3)

        def perturb(scale_deg, scale_t):
            P = np.eye(4)
            P[:3, :3] = D.ypr_to_R(*np.deg2rad(rng.uniform(-scale_deg, scale_deg, 3)))
            P[:3, 3] = rng.uniform(-scale_t, scale_t, 3)
            return T @ P

        sd, st = [(1, 0.05), (8, 0.05), (1, 0.4)][trial % 3]
        ops = [perturb(sd, st) for _ in range(3)]
        good = list(rng.uniform(0.3, 1.0, 3))
        dt = [3.0, 25.0][trial % 2]
        pf = [500, 1200][(trial // 2) % 2]
        c = ProcessedLoopCandidate(trial, 100.0 + dt, 100.0, 7, 3)
        c.opX_b_T_a, c.opX_goodness, c.pf_matches = ops, good, pf
        msg = c.makeLoopEdgeMsgWithConsistencyCheck()
        pub, pose, w = U.consistency_check(ops[0], ops[1], ops[2], good, dt, pf)
        assert (msg is not None) == pub
        if pub:
            n_pub += 1
            assert np.array_equal(msg.pose_1T0, pose) and abs(msg.weight - w) < 1e-15
            assert msg.description.startswith("7<=>3")
    assert n_pub > 10
    R = D.ypr_to_R(0.3, -0.2, 0.1)
    assert np.allclose(R2ypr(R), np.rad2deg([0.3, -0.2, 0.1]))
