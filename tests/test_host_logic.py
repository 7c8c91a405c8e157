"""Host-side logic that needs no GPU."""
import numpy as np
import pytest

from cerebro_b200.loop_detector import LoopEdge, ProcessedLoopCandidate, convert_channels  # noqa: F401


def test_channel_conversion_matches_opencv():
    """Cerebro.cpp:229-234 converts with cv::cvtColor; the mirror must give OpenCV's bytes."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (3, 37, 53, 3), dtype=np.uint8)
    bgr[0, 0, :6] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 2, 3]]
    gray = convert_channels(bgr, 1)
    assert gray.shape == (3, 37, 53, 1)
    for i in range(3):
        assert np.array_equal(gray[i, :, :, 0], cv2.cvtColor(bgr[i], cv2.COLOR_BGR2GRAY))
    g = rng.integers(0, 256, (2, 20, 31), dtype=np.uint8)
    rep = convert_channels(g, 3)
    assert rep.shape == (2, 20, 31, 3)
    for i in range(2):
        assert np.array_equal(rep[i], cv2.cvtColor(g[i], cv2.COLOR_GRAY2BGR))
    g14 = convert_channels(bgr, 1, fixed_point_bits=14)  # the OpenCV 3 table: at most one grey level away
    assert np.abs(g14.astype(int) - gray.astype(int)).max() <= 1 and (g14 != gray).mean() < 0.01
    assert convert_channels(bgr, 3) is not None and np.array_equal(convert_channels(bgr, 3), bgr)
    assert np.array_equal(convert_channels(g, 1)[..., 0], g)


def test_consistency_check_time_gate_is_floor_normalised():
    """ProcessedLoopCandidate.cpp:49-56 compares abs(ros::Duration::sec) with 10; sec is floor-normalised (nsec >= 0), so a
    difference of -9.5 s reads as sec = -10 and passes, while +9.5 s reads as 9 and is refused."""
    from cerebro_b200.loop_detector import ProcessedLoopCandidate

    def edge(t1, t2):
        c = ProcessedLoopCandidate(0, t1, t2, 7, 3)
        c.pf_matches = 900
        c.opX_b_T_a = [np.eye(4), np.eye(4), np.eye(4)]
        c.opX_goodness = [0.5, 0.7, 0.6]
        return c.makeLoopEdgeMsgWithConsistencyCheck()

    assert edge(100.0, 109.5) is not None  # t_1 - t_2 = -9.5 -> sec = -10
    assert edge(109.5, 100.0) is None      # +9.5 -> sec = 9
    assert edge(100.0, 109.0) is None      # -9.0 -> sec = -9
    e = edge(200.0, 100.0)
    assert e is not None and e.weight == 0.7 and e.description.startswith("7<=>3")


def test_descriptor_thread_dynamic_skip_rule():
    """Cerebro.cpp:189-203: after the first four keyframes, a keyframe is dropped when rand()/RAND_MAX <
    1 - incoming_diff_ms / estimated_descriptor_compute_time_ms; last_proc_timestamp advances either way."""
    from cerebro_b200.loop_detector import Cerebro

    c = Cerebro.__new__(Cerebro)
    c.estimated_descriptor_compute_time_ms, c._last_proc_timestamp, c._n_considered = 100, 0.0, 0
    seq = iter([0.05, 0.95, 0.2, 0.6])  # rand() is only drawn once n_computed > 4 (short-circuit, as in the reference)
    rand = lambda: int(next(seq) * Cerebro.RAND_MAX)
    stamps = [10.00, 10.02, 10.04, 10.06, 10.08, 10.10, 10.20, 10.23]  # 20 ms apart -> skip_frac 0.8; 100 ms -> 0; 30 -> 0.7
    got = [c._dynamic_skip(t, rand) for t in stamps]
    # first four are never skipped (n_computed > 4); then rand 0.05 < 0.8 skip, 0.95 keep, 100 ms gap: frac 0 keep, 0.6 < 0.7 skip
    assert got == [False, False, False, False, True, False, False, True]
    # a device-speed estimate (0 ms) never skips
    c.estimated_descriptor_compute_time_ms, c._n_considered = 0, 10
    assert not c._dynamic_skip(10.231, lambda: 0)


def test_cpp_consistency_check_and_loop_edge_match_python_mirror(native_lib, tmp_path):
    """C++ rendering of ProcessedLoopCandidate::makeLoopEdgeMsgWithConsistencyCheck + eigenmat_to_geometry_msgs_Pose
    (cerebro_b200/host/cerebro_shim.hpp, the code a ROS node keeps on the host) against the Python mirror on 900 trials around
    every threshold: |dt| = 10 s with the floor-normalised Duration, pf_matches = 800, 5 degree / 0.2 m deltas, and the
    reference's quirk (op1-icp translation tested twice, op1-op2 never).  Runs on CPU: no device call is involved."""
    import json
    import subprocess

    from cerebro_b200 import build
    from cerebro_b200.loop_detector import ProcessedLoopCandidate
    from oracle.dls_pnp import ypr_to_R

    rng = np.random.default_rng(17)

    def pose(ypr_deg, t):
        T = np.eye(4)
        T[:3, :3] = ypr_to_R(*np.deg2rad(ypr_deg))
        T[:3, 3] = t
        return T

    trials = []
    for i in range(900):
        base = pose(rng.uniform(-180, 180, 3) * [1, 0.45, 1], rng.uniform(-3, 3, 3))
        scale_r = [0.5, 2.0, 6.0, 12.0][i % 4]       # degrees: well inside, near, just outside, far outside 5
        scale_t = [0.02, 0.08, 0.25, 0.6][(i // 4) % 4]  # metres around 0.2
        ops = [base @ pose(rng.uniform(-1, 1, 3) * scale_r, rng.uniform(-1, 1, 3) * scale_t) for _ in range(3)]
        if i % 10 == 0:  # op1-op2 translation far apart, op1-icp and op2-icp close: only possible through the quirk -> stays rejected by op2-icp
            ops[1] = ops[0] @ pose([0, 0, 0], [0.5, 0, 0])
        t1 = float(rng.uniform(100, 200))
        dt = [-25.0, -10.0, -9.5, -9.999, 9.0, 9.999, 10.0, 10.4, 35.0][i % 9]
        pf = int([801, 800, 5000, 799, 2000][i % 5])
        g = rng.uniform(0, 1, 3)
        trials.append((t1, t1 - dt, pf, g, ops))
    path = str(tmp_path / "trials.raw")
    with open(path, "wb") as f:
        for t1, t2, pf, g, ops in trials:
            f.write(np.concatenate([[t1, t2, pf], g] + [o.reshape(16) for o in ops]).astype(np.float64).tobytes())
    r = subprocess.run([build.HARNESS, "--consistency", path, str(len(trials))], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = [json.loads(x) for x in r.stdout.strip().splitlines()]
    assert len(lines) == len(trials)
    n_ok = 0
    for i, ((t1, t2, pf, g, ops), out) in enumerate(zip(trials, lines)):
        t1q, t2q = round(t1 * 1e9) / 1e9, round(t2 * 1e9) / 1e9  # the harness keeps ros::Time's nanosecond grid
        c = ProcessedLoopCandidate(i, t1q, t2q, 2 * i, 2 * i + 1)
        c.pf_matches = pf
        c.opX_b_T_a = list(ops)
        c.opX_goodness = [float(np.float32(x)) for x in g]
        e = c.makeLoopEdgeMsgWithConsistencyCheck()
        assert bool(out["ok"]) == (e is not None), (i, out, t1 - t2, pf)
        if e is None:
            continue
        n_ok += 1
        assert out["description"] == e.description
        assert abs(out["weight"] - e.weight) < 1e-7
        assert np.allclose(out["position"], e.pose_1T0[:3, 3], atol=0, rtol=0)
        x, y, z, w = out["orientation"]
        assert abs(x * x + y * y + z * z + w * w - 1.0) < 1e-12
        Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(Rq, e.pose_1T0[:3, :3], atol=1e-12)
    assert 20 < n_ok < 450  # both outcomes are exercised
