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
