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
