"""Image side of a loop candidate behind the C ABI: the undistortion / stereo-rectification warps
(``StereoGeometry``, src/utils/CameraGeometry.cpp:42, 381-382: ``cv::remap(..., CV_INTER_LINEAR)``) and ORB extraction
(``StaticPointFeatureMatching::gms_point_feature_matches``, src/utils/PointFeatureMatching.cpp:16-22:
``cv::ORB::create(n_orb_feat)``, ``setFastThreshold(0)``, ``detectAndCompute``).  All arithmetic happens in
``libcerebro_b200.so`` (``cb_features_*``); there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


class Features:
    def __init__(self, rows: int, cols: int, max_images: int = 2, max_keypoints: int = 6000, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.rows, self.cols, self.max_images, self.max_keypoints = rows, cols, max_images, max_keypoints
        check(self._lib.cb_features_create(C.byref(self._h), rows, cols, max_images, max_keypoints, device))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.cb_features_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- StereoGeometry: maps are made once on the host (camodocal / cv::initUndistortRectifyMap), warps run on the device
    def set_remap(self, slot: int, map_x: np.ndarray, map_y: np.ndarray) -> None:
        mx = np.ascontiguousarray(map_x, dtype=np.float32)
        my = np.ascontiguousarray(map_y, dtype=np.float32)
        assert mx.shape == (self.rows, self.cols) and my.shape == (self.rows, self.cols)
        check(self._lib.cb_features_set_remap(self._h, slot, ptr(mx), ptr(my)))

    def remap(self, images_u8: np.ndarray, slot_a: int, slot_b: int = -1) -> np.ndarray:
        """``cv::remap(im, out, map_x, map_y, CV_INTER_LINEAR)`` for [n, rows, cols] images; a second slot chains the
        stereo-rectification warp after the undistortion warp (CameraGeometry.cpp:42 then :381-382)."""
        x = np.ascontiguousarray(images_u8, dtype=np.uint8)
        if x.ndim == 2:
            x = x[None]
        assert x.shape[1:] == (self.rows, self.cols)
        out = np.empty_like(x)
        check(self._lib.cb_features_remap(self._h, x.shape[0], ptr(x), slot_a, slot_b, ptr(out)))
        return out

    # ---- cv::ORB::create(n) + setFastThreshold(0) + detectAndCompute
    def orb(self, images_u8: np.ndarray, n_features: int = 5000):
        """[n, rows, cols] uint8 -> list of dicts (pt [k,2] float32, size, angle, response, octave, desc [k,32] uint8), one per
        image, keypoints in cv::ORB's output order."""
        x = np.ascontiguousarray(images_u8, dtype=np.uint8)
        if x.ndim == 2:
            x = x[None]
        assert x.shape[1:] == (self.rows, self.cols)
        n, m = x.shape[0], self.max_keypoints
        cnt = np.zeros(n, dtype=np.int32)
        xy = np.empty((n, m, 2), dtype=np.float32)
        size = np.empty((n, m), dtype=np.float32)
        ang = np.empty((n, m), dtype=np.float32)
        resp = np.empty((n, m), dtype=np.float32)
        octv = np.empty((n, m), dtype=np.int32)
        desc = np.empty((n, m, 32), dtype=np.uint8)
        check(self._lib.cb_features_orb(self._h, n, ptr(x), n_features, ptr(cnt), ptr(xy), ptr(size), ptr(ang), ptr(resp), ptr(octv), ptr(desc)))
        out = []
        for i in range(n):
            k = int(cnt[i])
            out.append(dict(pt=xy[i, :k].copy(), size=size[i, :k].copy(), angle=ang[i, :k].copy(), response=resp[i, :k].copy(),
                            octave=octv[i, :k].copy(), desc=desc[i, :k].copy()))
        return out

    def debug_read(self, what: int) -> np.ndarray:
        """Pyramid (0), FAST score map (1) or blurred pyramid (2) of image 0 of the last ``orb`` call, levels back to back."""
        buf = np.empty(4 * self.rows * self.cols, dtype=np.uint8)
        n = self._lib.cb_features_debug_read(self._h, what, ptr(buf), buf.size)
        if n < 0:
            check(int(n))
        return buf[:n].copy()

    @property
    def last_orb_ms(self) -> float:
        return float(self._lib.cb_features_last_orb_ms(self._h))
