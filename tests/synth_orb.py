"""Seeded inputs of the ORB / remap fixtures (shared by tools/make_golden_orb.py and tests/test_orb.py)."""
import numpy as np

from cerebro_b200 import synthetic

# name, rows, cols, n_features, image kind, seed
CASES = [("scene_480x752", 480, 752, 5000, "textured", 3), ("scene_480x640", 480, 640, 5000, "textured", 4),
         ("noise_240x320", 240, 320, 1500, "noise", 5), ("odd_197x263", 197, 263, 800, "textured", 6)]
REMAP_CASES = [("remap_480x752", 480, 752, 11), ("remap_197x263", 197, 263, 12)]


def image(kind, h, w, seed):
    f = synthetic.textured_scenes if kind == "textured" else synthetic.band_limited_images
    return np.ascontiguousarray(f(1, h, w, 1, seed=seed)[0, :, :, 0])


def rect_maps(h, w, seed):
    """A plausible undistortion / rectification map pair: small rotation + radial term, partly pointing outside the image."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2 + rng.uniform(-8, 8), h / 2 + rng.uniform(-8, 8)
    th = np.deg2rad(rng.uniform(-2, 2))
    xn, yn = (xx - cx) / w, (yy - cy) / w
    r2 = xn * xn + yn * yn
    k1 = rng.uniform(-0.25, 0.25)
    xd, yd = xn * (1 + k1 * r2), yn * (1 + k1 * r2)
    mx = (np.cos(th) * xd - np.sin(th) * yd) * w + cx + rng.uniform(-5, 5)
    my = (np.sin(th) * xd + np.cos(th) * yd) * w + cy + rng.uniform(-5, 5)
    return mx.astype(np.float32), my.astype(np.float32)
