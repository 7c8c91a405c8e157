"""Seeded synthetic rectified stereo pairs (numpy only, so the fixtures can be regenerated without OpenCV)."""
import numpy as np


def _blur(img, sigma):
    r = int(3 * sigma) + 1
    k = np.exp(-0.5 * (np.arange(-r, r + 1) / sigma) ** 2)
    k /= k.sum()
    img = np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 0, img)
    return np.apply_along_axis(lambda v: np.convolve(np.pad(v, r, mode="reflect"), k, mode="valid"), 1, img)


def stereo_scene(h, w, kind=0, seed=0, noise=1.5):
    """Textured scene seen by a rectified pair: left(x + d(x, y)) == right(x) up to noise; kind 0 = smooth disparity,
    1 = piecewise constant, 2 = ramp with a textureless band.  Returns (left, right) uint8 [h, w]."""
    rng = np.random.default_rng(seed)
    pad = 70
    tex = _blur(rng.random((h, w + 2 * pad)), 1.2)
    tex = (tex - tex.min()) / (tex.max() - tex.min()) * 255.0
    left = tex[:, pad : pad + w]
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    if kind == 0:
        d = 14 + 8 * np.sin(xs / 60) + 5 * np.cos(ys / 45)
    elif kind == 1:
        d = np.where(xs < w / 2, 6.0, 25.0) + np.where(ys < h / 3, 4.0, 0.0)
    else:
        d = 2 + 26 * (ys / h)
    src = xs + pad + d  # right(x) = tex(x + pad + d)
    x0 = np.floor(src).astype(int)
    a = src - x0
    rows = np.arange(h)[:, None]
    right = (1 - a) * tex[rows, np.clip(x0, 0, tex.shape[1] - 1)] + a * tex[rows, np.clip(x0 + 1, 0, tex.shape[1] - 1)]
    right = right + rng.normal(0, noise, (h, w))
    left = np.clip(np.rint(left), 0, 255).astype(np.uint8)
    right = np.clip(np.rint(right), 0, 255).astype(np.uint8)
    if kind == 2:
        left[h // 2 - 12 : h // 2 + 12] = 128
        right[h // 2 - 12 : h // 2 + 12] = 128
    return left, right
