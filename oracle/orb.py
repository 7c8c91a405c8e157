"""ORACLE (test infrastructure, not product): numpy restatement of ``cv::ORB::create(n)`` + ``setFastThreshold(0)`` +
``detectAndCompute`` as the reference calls it (src/utils/PointFeatureMatching.cpp:16-22), following OpenCV's
modules/features2d/src/{orb,fast,fast_score,keypoint}.cpp and imgproc's bit-exact INTER_LINEAR_EXACT resize.

PINNED: bit-exact (keypoint order, coordinates, size, angle, response, octave and the 256-bit descriptors) against the
installed OpenCV -- the reference's own dependency -- in tests/test_orb.py, and on committed fixtures
(tests/golden/orb_golden.npz).  cv::ORB's keypoint order depends on the permutation libstdc++'s std::nth_element leaves
behind (KeyPointsFilter::retainBest); oracle/orb_select.cpp calls the same library routine on bare records.

Only tests/ may import this module.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EDGE_THRESHOLD = 31
PATCH_SIZE = 31
HALF_PATCH = 15
HARRIS_BLOCK = 7
HARRIS_K = np.float32(0.04)
N_LEVELS = 8
SCALE_FACTOR = float(np.float32(1.2))  # ORB::create takes a float; stored as double
BORDER = 32  # max(edgeThreshold, ceil(halfPatch * sqrt 2), HARRIS_BLOCK_SIZE / 2) + 1

_sel = None


def _select_lib():
    global _sel
    if _sel is None:
        _sel = ctypes.CDLL(os.path.join(HERE, "_ref", "liborb_select.so"))
        _sel.orb_retain_best.restype = ctypes.c_int
        _sel.orb_retain_best.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return _sel


def retain_best(response: np.ndarray, n_points: int) -> np.ndarray:
    """KeyPointsFilter::retainBest: indices (into the input order) of the survivors, in cv's output order."""
    r = np.ascontiguousarray(response, dtype=np.float32)
    out = np.empty(max(len(r), 1), dtype=np.int32)
    n = _select_lib().orb_retain_best(r.ctypes.data, len(r), int(n_points), out.ctypes.data)
    return out[:n].copy()


# ---------------------------------------------------------------------------------------------------------------------
# INTER_LINEAR_EXACT (imgproc/src/resize.cpp, resize_bitExact<uint8_t, interpolationLinear>): 8.8 fixed-point coefficients
# from softdouble positions, horizontal pass into 8.8 values, vertical pass in 16.16, round half up
# ---------------------------------------------------------------------------------------------------------------------
def _lin_coeffs(src: int, dst: int):
    inv = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv
    ofs = np.zeros(dst, np.int64)
    c1 = np.zeros(dst, np.int64)
    lo, hi = 0, dst
    for v in range(dst):
        fval = scale * (np.float64(v) + 0.5) - 0.5
        ival = math.floor(fval)
        if ival >= 0 and src > 1:
            if ival < src - 1:
                ofs[v] = ival
                c1[v] = int(np.rint((fval - ival) * 256.0))
            else:
                ofs[v] = src - 1
                hi = min(hi, v)
        else:
            lo = max(lo, v + 1)
    return ofs, c1, lo, hi


def resize_linear_exact(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    H, W = img.shape
    ox, cx, minx, maxx = _lin_coeffs(W, dw)
    oy, cy, miny, maxy = _lin_coeffs(H, dh)
    src = img.astype(np.int64)
    x = np.arange(dw)
    o1 = np.minimum(ox + 1, W - 1)
    hor = (256 - cx)[None, :] * src[:, ox] + cx[None, :] * src[:, o1]
    hor[:, x < minx] = src[:, :1] * 256
    hor[:, x >= maxx] = src[:, W - 1 :] * 256
    y = np.arange(dh)
    o1y = np.minimum(oy + 1, H - 1)
    v = (256 - cy)[:, None] * hor[oy] + cy[:, None] * hor[o1y]
    out = np.clip((v + (1 << 15)) >> 16, 0, 255)
    edge_lo, edge_hi = y < miny, y >= maxy
    out[edge_lo] = (hor[0][None, :] + 128) >> 8
    out[edge_hi] = (hor[H - 1][None, :] + 128) >> 8
    return out.astype(np.uint8)


def level_scales(n_levels: int = N_LEVELS):
    return [np.float32(math.pow(SCALE_FACTOR, lv)) for lv in range(n_levels)]


def build_pyramid(img: np.ndarray, n_levels: int = N_LEVELS):
    """Level images (each resized from the PREVIOUS level, orb.cpp) without their borders."""
    H, W = img.shape
    levels = [img]
    for lv in range(1, n_levels):
        sc = level_scales(n_levels)[lv]
        dw = int(np.rint(np.float32(W) / sc))
        dh = int(np.rint(np.float32(H) / sc))
        levels.append(resize_linear_exact(levels[-1], dw, dh))
    return levels


def features_per_level(nfeatures: int, n_levels: int = N_LEVELS):
    factor = np.float32(1.0 / SCALE_FACTOR)
    nd = np.float32(nfeatures * (1 - factor) / (1 - np.float32(math.pow(float(factor), n_levels))))
    out, total = [], 0
    for lv in range(n_levels - 1):
        out.append(int(np.rint(nd)))
        total += out[-1]
        nd = np.float32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# FAST-9/16 with threshold t and non-maximum suppression (features2d/src/fast.cpp, fast_score.cpp)
# ---------------------------------------------------------------------------------------------------------------------
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def fast_score_map(img: np.ndarray, threshold: int = 0) -> np.ndarray:
    """cornerScore<16> where the pixel is a FAST-9 corner at `threshold`, else 0 (what fast.cpp keeps in its row buffers)."""
    H, W = img.shape
    v = img.astype(np.int32)
    c = v[3 : H - 3, 3 : W - 3]
    d = np.stack([c - v[3 + dy : H - 3 + dy, 3 + dx : W - 3 + dx] for dx, dy in CIRCLE])  # d[k] = v - p_k
    best = np.full(c.shape, -(1 << 20), np.int32)
    for k in range(16):
        idx = [(k + i) % 16 for i in range(9)]
        a = d[idx]
        best = np.maximum(best, np.maximum(a.min(0), (-a).min(0)))
    score = np.zeros((H, W), np.int32)
    inner = np.where(best > threshold, best - 1, 0)
    score[3 : H - 3, 3 : W - 3] = inner
    return score


def fast_detect(img: np.ndarray, threshold: int = 0):
    """Keypoints (x, y, response) in cv::FAST's output order (row-major) after non-maximum suppression."""
    H, W = img.shape
    s = fast_score_map(img, threshold)
    is_corner = np.zeros((H, W), bool)
    v = img.astype(np.int32)
    # a corner with score 0 still occupies its slot but can never win the strict comparison
    c = s[1 : H - 1, 1 : W - 1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= c > s[1 + dy : H - 1 + dy, 1 + dx : W - 1 + dx]
    is_corner[1 : H - 1, 1 : W - 1] = keep
    # the corner test of the LAST processed row (H - 4) is never emitted: fast.cpp emits row i - 1 while processing row i
    ys, xs = np.nonzero(is_corner)
    return xs.astype(np.float32), ys.astype(np.float32), s[ys, xs].astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# Harris response, orientation, descriptors (features2d/src/orb.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def harris_responses(img: np.ndarray, xs, ys, block: int = HARRIS_BLOCK):
    v = img.astype(np.int32)
    r = block // 2
    out = np.empty(len(xs), np.float32)
    scale = np.float32(1.0) / np.float32((1 << 2) * block * np.float32(255.0))
    s4 = np.float32(np.float32(np.float32(scale * scale) * scale) * scale)
    Ix = (v[1:-1, 2:] - v[1:-1, :-2]) * 2 + (v[:-2, 2:] - v[:-2, :-2]) + (v[2:, 2:] - v[2:, :-2])
    Iy = (v[2:, 1:-1] - v[:-2, 1:-1]) * 2 + (v[2:, :-2] - v[:-2, :-2]) + (v[2:, 2:] - v[:-2, 2:])
    Ixx, Iyy, Ixy = Ix * Ix, Iy * Iy, Ix * Iy  # index [y-1, x-1]
    for i in range(len(xs)):
        x0, y0 = int(np.rint(xs[i])), int(np.rint(ys[i]))
        sl = (slice(y0 - r - 1, y0 + r), slice(x0 - r - 1, x0 + r))
        a, b, c = int(Ixx[sl].sum()), int(Iyy[sl].sum()), int(Ixy[sl].sum())
        fa, fb, fc = np.float32(a), np.float32(b), np.float32(c)
        t1 = np.float32(fa * fb)
        t2 = np.float32(fc * fc)
        sab = np.float32(fa + fb)
        t3 = np.float32(np.float32(HARRIS_K * sab) * sab)
        out[i] = np.float32(np.float32(np.float32(t1 - t2) - t3) * s4)
    return out


def _umax():
    umax = [0] * (HALF_PATCH + 2)
    vmax = int(math.floor(np.float32(HALF_PATCH) * np.float32(math.sqrt(2.0)) / 2 + 1))
    vmin = int(math.ceil(np.float32(HALF_PATCH) * np.float32(math.sqrt(2.0)) / 2))
    for v in range(vmax + 1):
        umax[v] = int(np.rint(math.sqrt(float(HALF_PATCH * HALF_PATCH - v * v))))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


UMAX = _umax()
_R2D = np.float32(180.0 / math.pi)  # `0.99...f*(float)(180/CV_PI)`: a float x float product, rounded once more
_P1 = np.float32(np.float32(0.9997878412794807) * _R2D)
_P3 = np.float32(np.float32(-0.3258083974640975) * _R2D)
_P5 = np.float32(np.float32(0.1555786518463281) * _R2D)
_P7 = np.float32(np.float32(-0.04432655554792128) * _R2D)
_EPS = np.float32(2.220446049250313e-16)


def fast_atan2(y, x, fma: bool = False):
    """cv::fastAtan2 (core/src/mathfuncs_core.simd.hpp atan_f32), degrees."""
    y, x = np.float32(y), np.float32(x)
    ax, ay = np.float32(abs(x)), np.float32(abs(y))

    def poly(c):
        c2 = np.float32(c * c)
        if fma:
            f = lambda a, b, cc: np.float32(np.float64(a) * np.float64(b) + np.float64(cc))
            return np.float32(f(f(f(_P7, c2, _P5), c2, _P3), c2, _P1) * c)
        return np.float32(np.float32(np.float32(np.float32(np.float32(np.float32(np.float32(_P7 * c2) + _P5) * c2) + _P3) * c2) + _P1) * c)

    if ax >= ay:
        a = poly(np.float32(ay / np.float32(ax + _EPS)))
    else:
        a = np.float32(np.float32(90.0) - poly(np.float32(ax / np.float32(ay + _EPS))))
    if x < 0:
        a = np.float32(np.float32(180.0) - a)
    if y < 0:
        a = np.float32(np.float32(360.0) - a)
    return a


def ic_angles(img_padded: np.ndarray, xs, ys, fma: bool = False):
    """ICAngles on the level image padded by BORDER (reflect-101), keypoints in level coordinates."""
    v = img_padded.astype(np.int64)
    out = np.empty(len(xs), np.float32)
    us = np.arange(-HALF_PATCH, HALF_PATCH + 1)
    for i in range(len(xs)):
        cx, cy = int(np.rint(xs[i])) + BORDER, int(np.rint(ys[i])) + BORDER
        m10 = int((us * v[cy, cx - HALF_PATCH : cx + HALF_PATCH + 1]).sum())
        m01 = 0
        for dv in range(1, HALF_PATCH + 1):
            d = UMAX[dv]
            u = np.arange(-d, d + 1)
            plus, minus = v[cy + dv, cx - d : cx + d + 1], v[cy - dv, cx - d : cx + d + 1]
            m01 += dv * int((plus - minus).sum())
            m10 += int((u * (plus + minus)).sum())
        out[i] = fast_atan2(m01, m10, fma)
    return out


# The learned 256 x 2 x 2 test-point pattern of rBRIEF (`bit_pattern_31_` in features2d/src/orb.cpp; Rublee et al., "ORB: an
# efficient alternative to SIFT or SURF", ICCV 2011) is DATA, shipped with the product table (cerebro_b200/csrc/orb_pattern.h);
# the oracle keeps its own copy in tests/golden/orb_pattern.npy.
def load_pattern():
    p = os.path.join(os.path.dirname(HERE), "tests", "golden", "orb_pattern.npy")
    return np.load(p).astype(np.int32).reshape(512, 2)


GAUSS7 = np.array([18, 34, 48, 56, 48, 34, 18], np.int64)  # sigma 2, 8 fractional bits, error-diffused to sum 256


def gaussian_blur7(img: np.ndarray, kernel=GAUSS7) -> np.ndarray:
    """GaussianBlur(Size(7,7), 2, 2, BORDER_REFLECT_101) on 8-bit data: separable fixed point, 16.16 accumulate, round half up."""
    H, W = img.shape
    p = np.pad(img.astype(np.int64), 3, mode="reflect")
    hor = sum(kernel[i] * p[:, i : i + W] for i in range(7))
    ver = sum(kernel[i] * hor[i : i + H, :] for i in range(7))
    return np.clip((ver + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


def compute_descriptors(blurred_padded_levels, kps, pattern):
    """computeOrbDescriptors (WTA_K = 2).  kps: rows (x, y, size, angle, response, octave) in level-0 coordinates;
    blurred_padded_levels[L] = blurred level image padded by BORDER (reflect-101 of the UNBLURRED image outside)."""
    scales = level_scales()
    out = np.zeros((len(kps), 32), np.uint8)
    d2r = np.float32(math.pi / np.float32(180.0))
    px, py = pattern[:, 0].astype(np.float32), pattern[:, 1].astype(np.float32)
    for j, (x, y, _size, angle, _resp, octv) in enumerate(kps):
        L = int(octv)
        img = blurred_padded_levels[L]
        inv = np.float32(np.float32(1.0) / scales[L])
        ang = np.float32(np.float32(angle) * d2r)
        a, b = np.float32(math.cos(float(ang))), np.float32(math.sin(float(ang)))
        cy = int(np.rint(np.float32(np.float32(y) * inv))) + BORDER
        cx = int(np.rint(np.float32(np.float32(x) * inv))) + BORDER
        rx = np.rint((px * a).astype(np.float32) - (py * b).astype(np.float32)).astype(np.int64)
        ry = np.rint((px * b).astype(np.float32) + (py * a).astype(np.float32)).astype(np.int64)
        v = img[cy + ry, cx + rx].astype(np.int32)
        bits = (v[0::2] < v[1::2]).astype(np.uint8)  # 256 comparisons
        out[j] = np.packbits(bits.reshape(32, 8), axis=1, bitorder="little")[:, 0]
    return out


def blur_float(img: np.ndarray) -> np.ndarray:
    """The 7x7 sigma-2 Gaussian as cv::ORB actually applies it (a pyramid ROI is a sub-matrix, so OpenCV leaves its fixed-point
    branch and runs a FLOAT separable filter -- IPP's in the installed build): row pass then column pass in float32 with the
    taps of getGaussianKernel(7, 2, CV_32F), round to nearest even."""
    k = np.array([0.07015932, 0.13107488, 0.19071282, 0.21610594, 0.19071282, 0.13107488, 0.07015932], np.float32)
    H, W = img.shape
    p = np.pad(img.astype(np.float32), 3, mode="reflect")
    hor = (k[0] * p[:, 0:W]).astype(np.float32)
    for i in range(1, 7):
        hor = (hor + (k[i] * p[:, i : i + W]).astype(np.float32)).astype(np.float32)
    ver = (k[0] * hor[0:H]).astype(np.float32)
    for i in range(1, 7):
        ver = (ver + (k[i] * hor[i : i + H]).astype(np.float32)).astype(np.float32)
    return np.clip(np.rint(ver), 0, 255).astype(np.uint8)


def detect_and_compute(img: np.ndarray, nfeatures: int = 5000):
    """cv::ORB::create(nfeatures) + setFastThreshold(0) + detectAndCompute.  Returns (kps [n, 6] float64 rows of
    (x, y, size, angle, response, octave) in cv's output order, descriptors [n, 32] uint8)."""
    lv = build_pyramid(img)
    npl = features_per_level(nfeatures)
    scales = level_scales()
    stage1 = []
    for L in range(N_LEVELS):
        h, w = lv[L].shape
        xs, ys, rs = fast_detect(lv[L], 0)
        inb = (xs >= EDGE_THRESHOLD) & (xs < w - EDGE_THRESHOLD) & (ys >= EDGE_THRESHOLD) & (ys < h - EDGE_THRESHOLD)
        xs, ys, rs = xs[inb], ys[inb], rs[inb]
        keep = retain_best(rs, 2 * npl[L])
        stage1.append((xs[keep], ys[keep]))
    rows = []
    for L in range(N_LEVELS):
        xs, ys = stage1[L]
        hr = harris_responses(lv[L], xs, ys)
        keep = retain_best(hr, npl[L])
        xs, ys, hr = xs[keep], ys[keep], hr[keep]
        ang = ic_angles(np.pad(lv[L], BORDER, mode="reflect"), xs, ys)
        sc = scales[L]
        for i in range(len(xs)):
            rows.append((np.float32(xs[i] * sc), np.float32(ys[i] * sc), np.float32(np.float32(PATCH_SIZE) * sc), ang[i], hr[i], L))
    kps = np.array(rows, np.float64).reshape(-1, 6)
    blurred = []
    for L in range(N_LEVELS):
        pad = np.pad(lv[L], BORDER, mode="reflect")
        pad[BORDER:-BORDER, BORDER:-BORDER] = blur_float(lv[L])
        blurred.append(pad)
    desc = compute_descriptors(blurred, [tuple(r) for r in kps], load_pattern())
    return kps, desc
