"""ORACLE (test infrastructure, not product): the stereo depth step of the geometric verifier's front end.

Restates, with file:line,
  * ``bm = cv::StereoBM::create(64, 21); bm->compute(left, right, disparity)``
        src/utils/CameraGeometry.cpp:81, :410-418 (``do_stereoblockmatching_of_srectified_images``) -- OpenCV's block
        matcher with its default parameters (XSOBEL pre-filter cap 31, texture threshold 10, uniqueness ratio 15, no
        speckle filter, no left-right check): pre-filter, 21 x 21 SAD over 64 disparities with replicated borders, first
        minimum, texture / uniqueness rejection, parabola sub-pixel fit, disparity * 16 as int16, FILTERED = -16 outside
        the valid ROI.  OpenCV is a third-party dependency of the reference (not vendored, no version pin; README lists
        3.x); the algorithm below follows modules/calib3d/src/stereobm.cpp.
  * ``StereoGeometry::disparity_to_3DPoints``   src/utils/CameraGeometry.cpp:459-520 (the reference's own loop, not
        cv::reprojectImageTo3D):  pw = 1.0f / (disp / 16. * Q32 + Q33 + 1e-6);  X = ((j + Q03) pw, (i + Q13) pw, Q23 pw).

PINNED: ``stereo_bm`` is bit-exact against the installed OpenCV's ``cv2.StereoBM_create(ndisp, wsz).compute`` on every
case of tests/test_stereo.py (sizes incl. odd heights, 16..128 disparities, windows 5..21) and against the fixture that
cv2 produced (tests/golden/stereo_golden.npz, tools/make_golden_stereo.py).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

PREFILTER_CAP = 31  # StereoBM default preFilterCap
TEXTURE_THRESHOLD = 10
UNIQUENESS_RATIO = 15


def prefilter_xsobel(src: np.ndarray, ftzero: int = PREFILTER_CAP) -> np.ndarray:
    """stereobm.cpp prefilterXSobel: clamp(dx(y-1) + 2 dx(y) + dx(y+1), -cap, cap) + cap with dx = I[x+1] - I[x-1].
    Rows are produced in pairs: the row above the first row is row 1, the row below the last row of a pair that ends the
    image is the pair's FIRST row, and the last row of an odd-height image, like the first and last column, is ``cap``."""
    src = np.asarray(src, dtype=np.uint8)
    h, w = src.shape
    s = src.astype(np.int32)
    dx = np.zeros((h, w), dtype=np.int32)
    dx[:, 1:-1] = s[:, 2:] - s[:, :-2]
    out = np.full((h, w), ftzero, dtype=np.uint8)

    def tab(v):
        return np.clip(v, -ftzero, ftzero) + ftzero

    y = 0
    while y < h - 1:
        r0 = y - 1 if y > 0 else y + 1
        r3 = y + 2 if y < h - 2 else y
        out[y, 1:-1] = tab(dx[r0] + 2 * dx[y] + dx[y + 1])[1:-1]
        out[y + 1, 1:-1] = tab(dx[y] + 2 * dx[y + 1] + dx[r3])[1:-1]
        y += 2
    return out


def _box_rows_replicate(a: np.ndarray, r: int) -> np.ndarray:
    """Sum over rows y-r..y+r with the row index clamped to the image (findStereoCorrespondenceBM's hsad / htext borders)."""
    h = a.shape[0]
    p = a[np.clip(np.arange(-r, h + r), 0, h - 1)]
    c = np.concatenate([np.zeros((1,) + p.shape[1:], dtype=p.dtype), np.cumsum(p, axis=0)], axis=0)
    return c[2 * r + 1 :] - c[: -(2 * r + 1)]


def stereo_bm(left: np.ndarray, right: np.ndarray, ndisp: int = 64, wsz: int = 21, mindisp: int = 0, ftzero: int = PREFILTER_CAP,
              texture_threshold: int = TEXTURE_THRESHOLD, uniqueness_ratio: int = UNIQUENESS_RATIO) -> np.ndarray:
    """uint8 [h, w] rectified left / right -> int16 [h, w] disparity * 16 (FILTERED = (mindisp - 1) * 16)."""
    L = prefilter_xsobel(left, ftzero).astype(np.int32)
    R = prefilter_xsobel(right, ftzero).astype(np.int32)
    h, w = L.shape
    wsz2 = wsz // 2
    lofs = max(ndisp - 1 + mindisp, 0)
    rofs = -min(ndisp - 1 + mindisp, 0)
    width1 = w - rofs - ndisp + 1
    FILTERED = (mindisp - 1) << 4
    disp = np.full((h, w), FILTERED, dtype=np.int16)
    if lofs >= w or rofs >= w or width1 < 1:
        return disp
    # output column x (image column lofs + x) sums window columns c = x - wsz2 .. x + wsz2, clamped per image
    c = np.arange(-wsz2, width1 + wsz2)
    lcol = lofs + np.clip(c, -lofs, w - lofs - 1)
    rcol = rofs + np.clip(c, -rofs, w - rofs - ndisp)
    Lc = L[:, lcol]
    sad = np.empty((ndisp, h, width1), dtype=np.int32)
    for d in range(ndisp):  # candidate d compares left column X with right column X - (ndisp - 1 - d)
        diff = np.abs(Lc - R[:, rcol + d])
        cs = np.concatenate([np.zeros((h, 1), dtype=np.int64), np.cumsum(diff, axis=1)], axis=1)
        sad[d] = _box_rows_replicate(cs[:, wsz:] - cs[:, :-wsz], wsz2)
    text = np.abs(Lc - ftzero)
    cs = np.concatenate([np.zeros((h, 1), dtype=np.int64), np.cumsum(text, axis=1)], axis=1)
    tsum = _box_rows_replicate(cs[:, wsz:] - cs[:, :-wsz], wsz2)
    mind = sad.argmin(axis=0)  # strict '<' while d ascends: the first minimum
    minsad = np.take_along_axis(sad, mind[None], 0)[0]
    ok = tsum >= texture_threshold
    if uniqueness_ratio > 0:
        thresh = minsad + (minsad * uniqueness_ratio) // 100
        dd = np.arange(ndisp)[:, None, None]
        far = (dd < mind[None] - 1) | (dd > mind[None] + 1)
        ok &= ~(far & (sad <= thresh[None])).any(axis=0)
    ext = np.concatenate([sad[1:2], sad, sad[ndisp - 2 : ndisp - 1]], axis=0)  # sad[-1] = sad[1], sad[ndisp] = sad[ndisp-2]
    p = np.take_along_axis(ext, (mind + 2)[None], 0)[0].astype(np.int64)
    n = np.take_along_axis(ext, mind[None], 0)[0].astype(np.int64)
    den = p + n - 2 * minsad.astype(np.int64) + np.abs(p - n)
    q = np.where(den != 0, np.trunc((p - n) * 256 / np.where(den == 0, 1, den)).astype(np.int64), 0)  # C '/' truncates
    val = ((ndisp - mind - 1 + mindisp) * 256 + q + 15) >> 4  # dispDescale<short>
    disp[:, lofs : lofs + width1] = np.where(ok, val, FILTERED).astype(np.int16)
    # getValidDisparityROI with the default (empty) roi1 / roi2: everything outside it is FILTERED
    xmin, xmax = max(0, mindisp + ndisp - 1) + wsz2, min(w, w - mindisp) - wsz2
    ymin, ymax = wsz2, h - wsz2
    if xmax - xmin > 0 and ymax - ymin > 0:
        crop = np.full_like(disp, FILTERED)
        crop[ymin:ymax, xmin:xmax] = disp[ymin:ymax, xmin:xmax]
        disp = crop
    return disp


def disparity_to_3d(disparity_raw: np.ndarray, Q03: float, Q13: float, Q23: float, Q32: float, Q33: float) -> np.ndarray:
    """StereoGeometry::disparity_to_3DPoints, CameraGeometry.cpp:500-520: int16 disparity * 16 -> float32 [h, w, 3].
    The Q entries are floats (:486-499); ``disp / 16.`` and the denominator are evaluated in double, ``1.0f / denom`` is
    rounded to float, the three products are float."""
    d = np.asarray(disparity_raw).astype(np.float32)
    h, w = d.shape
    q03, q13, q23, q32, q33 = (np.float32(v) for v in (Q03, Q13, Q23, Q32, Q33))
    denom = d.astype(np.float64) / 16.0 * np.float64(q32) + np.float64(q33) + 1e-6
    pw = (1.0 / denom).astype(np.float32)
    j = np.arange(w, dtype=np.float32)[None, :]
    i = np.arange(h, dtype=np.float32)[:, None]
    out = np.empty((h, w, 3), dtype=np.float32)
    out[..., 0] = (j + q03) * pw
    out[..., 1] = (i + q13) * pw
    out[..., 2] = q23 * pw
    return out
