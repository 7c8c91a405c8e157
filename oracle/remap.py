"""ORACLE (test infrastructure, not product): numpy restatement of ``cv::remap(src, dst, map_x, map_y, CV_INTER_LINEAR)`` with
CV_32FC1 maps and the default constant border, as the reference applies it for undistortion and stereo rectification
(src/utils/CameraGeometry.cpp:42, 381-382).  Follows imgproc/src/imgwarp.cpp: map coordinates are rounded to 1/32 pixel
(``cvRound(x * INTER_TAB_SIZE)``), the four bilinear weights are integers summing to 2^15 (exact for a 32 x 32 table), the
result is ``(sum + 2^14) >> 15``; taps outside the image read the border value 0.

PINNED: bit-exact against the installed OpenCV (tests/test_orb.py).  Only tests/ may import this module."""
import numpy as np


def remap_linear(img: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    H, W = img.shape
    sx = np.rint(map_x.astype(np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(map_y.astype(np.float32) * np.float32(32)).astype(np.int64)
    fx, fy = sx & 31, sy & 31
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0).astype(np.int64)

    w00, w01 = (32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32
    w10, w11 = fy * (32 - fx) * 32, fy * fx * 32
    v = w00 * tap(iy, ix) + w01 * tap(iy, ix + 1) + w10 * tap(iy + 1, ix) + w11 * tap(iy + 1, ix + 1)
    return ((v + (1 << 14)) >> 15).astype(np.uint8)
