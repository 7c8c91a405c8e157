"""Host-side mirror of the correspondence front end, ``StaticPointFeatureMatching``
(src/utils/PointFeatureMatching.cpp), over ``cb_frontend_*``:

  gms_point_feature_matches        :5-75    ORB (OpenCV, host) -> BFMatcher(NORM_HAMMING).match -> gms_matcher -> u, ud
  make_3d_2d_collection__using__pfmatches_and_disparity   :96-153
  make_3d_3d_collection__using__pfmatches_and_disparity   :158-196

The matcher, the GMS filter and the set builders run on the device for a whole batch of loop candidates at once; ORB
detection / description and the stereo block matcher remain OpenCV calls on the host -- the
batch entry points therefore take keypoints + descriptors, and ``gms_point_feature_matches`` (images in, like the
reference) is a convenience that runs cv2.ORB first when OpenCV's Python module is importable.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, ptr


class FrontEnd:
    def __init__(self, max_pairs: int = 16, max_features: int = 5000, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.cb_frontend_create(C.byref(self._h), max_pairs, max_features, device))
        self.max_pairs, self.max_features = max_pairs, max_features
        self._off1 = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.cb_frontend_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- BFMatcher(NORM_HAMMING).match + gms_matcher(...).GetInlierMask(mask, false, false), batched
    def match_gms(self, kp1: Sequence[np.ndarray], desc1: Sequence[np.ndarray], kp2: Sequence[np.ndarray],
                  desc2: Sequence[np.ndarray], size1: Tuple[int, int], size2: Tuple[int, int]):
        """Per pair p: kp1[p] [n1,2] float32 KeyPoint.pt, desc1[p] [n1,32] uint8 (query = frame a), kp2 / desc2 (train =
        frame b); size = (width, height).  Returns a list of dicts with ``train_idx``, ``distance``, ``inliers`` (bool mask),
        ``n_inliers`` -- one match per query descriptor, in query order, exactly what ``matcher.match(d1, d2)`` returns."""
        n = len(kp1)
        assert n == len(desc1) == len(kp2) == len(desc2) and n >= 1
        off1 = np.zeros(n + 1, dtype=np.int32)
        off2 = np.zeros(n + 1, dtype=np.int32)
        off1[1:] = np.cumsum([len(k) for k in kp1])
        off2[1:] = np.cumsum([len(k) for k in kp2])
        cat = lambda xs, dt, w: (np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=dt).reshape(-1, w) for x in xs]))
                                 if sum(len(x) for x in xs) else np.zeros((0, w), dtype=dt))
        K1, K2 = cat(kp1, np.float32, 2), cat(kp2, np.float32, 2)
        D1, D2 = cat(desc1, np.uint8, 32), cat(desc2, np.uint8, 32)
        assert K1.shape[0] == D1.shape[0] == off1[-1] and K2.shape[0] == D2.shape[0] == off2[-1]
        tot = max(int(off1[-1]), 1)
        tidx = np.full(tot, -1, dtype=np.int32)
        dist = np.full(tot, -1, dtype=np.int32)
        mask = np.zeros(tot, dtype=np.uint8)
        ninl = np.zeros(n, dtype=np.int32)
        # zero-length inputs still need a valid pointer
        pad = np.zeros(32, dtype=np.uint8)
        check(self._lib.cb_frontend_match_gms(self._h, n, ptr(off1), ptr(off2), ptr(D1 if D1.size else pad), ptr(D2 if D2.size else pad),
                                              ptr(K1 if K1.size else pad), ptr(K2 if K2.size else pad), int(size1[0]), int(size1[1]),
                                              int(size2[0]), int(size2[1]), ptr(tidx), ptr(dist), ptr(mask), ptr(ninl)))
        self._off1, self._K1, self._K2, self._off2 = off1, K1, K2, off2
        self._tidx, self._mask = tidx, mask
        out = []
        for p in range(n):
            a, b = off1[p], off1[p + 1]
            out.append(dict(train_idx=tidx[a:b].copy(), distance=dist[a:b].copy(), inliers=mask[a:b].astype(bool), n_inliers=int(ninl[p])))
        return out

    def last_match_ms(self) -> float:
        return float(self._lib.cb_frontend_last_match_ms(self._h))

    def matched_points(self, p: int):
        """``MiscUtils::dmatch_2_eigen(kp1, kp2, matches_gms, u, ud, true)`` (PointFeatureMatching.cpp:71): the GMS inliers
        of pair p as 3xN homogeneous float64 pixel coordinates (u of frame a, ud of frame b)."""
        a, b = self._off1[p], self._off1[p + 1]
        sel = np.nonzero(self._mask[a:b])[0]
        u = self._K1[a:b][sel].astype(np.float64)
        ud = self._K2[self._off2[p] : self._off2[p + 1]][self._tidx[a:b][sel]].astype(np.float64)
        one = np.ones((1, sel.size))
        return np.concatenate([u.T, one]), np.concatenate([ud.T, one])

    # ---- stereo depth of the frames (StereoGeometry, src/utils/CameraGeometry.cpp:81, :410-418, :459-520)
    def stereo_bm(self, left: np.ndarray, right: np.ndarray, ndisp: int = 64, wsz: int = 21) -> np.ndarray:
        """``cv::StereoBM::create(ndisp, wsz)->compute(left, right)`` for a batch of rectified pairs: uint8 [n, rows, cols]
        (or [rows, cols]) -> int16 disparity * 16 of the same shape, -16 where rejected."""
        single = left.ndim == 2
        L = np.ascontiguousarray(left[None] if single else left, dtype=np.uint8)
        R = np.ascontiguousarray(right[None] if single else right, dtype=np.uint8)
        assert L.ndim == 3 and L.shape == R.shape
        out = np.empty(L.shape, dtype=np.int16)
        check(self._lib.cb_frontend_stereo_bm(self._h, L.shape[0], ptr(L), ptr(R), L.shape[1], L.shape[2], int(ndisp), int(wsz), ptr(out)))
        return out[0] if single else out

    def last_stereo_ms(self) -> float:
        return float(self._lib.cb_frontend_last_stereo_ms(self._h))

    def disparity_to_3d(self, disparity: np.ndarray, Q: np.ndarray) -> np.ndarray:
        """``StereoGeometry::disparity_to_3DPoints``: int16 disparity * 16 [n, rows, cols] (or [rows, cols]) and the 4x4
        reprojection matrix Q -> the CV_32FC3 "3d image" float32 [..., 3] that the set builders take."""
        single = disparity.ndim == 2
        D = np.ascontiguousarray(disparity[None] if single else disparity, dtype=np.int16)
        Q = np.asarray(Q, dtype=np.float64)
        out = np.empty(D.shape + (3,), dtype=np.float32)
        check(self._lib.cb_frontend_disparity_to_3d(self._h, D.shape[0], ptr(D), D.shape[1], D.shape[2], float(Q[0, 3]), float(Q[1, 3]),
                                                    float(Q[2, 3]), float(Q[3, 2]), float(Q[3, 3]), ptr(out)))
        return out[0] if single else out

    # ---- set builders over the batch matched last
    def make_3d_2d_collection(self, K: np.ndarray, img3d: np.ndarray, swapped: bool = False):
        """make_3d_2d_collection__using__pfmatches_and_disparity for every pair of the last ``match_gms`` batch.
        img3d [n_pairs, H, W, 3] float32 = the 3-D image of frame a -- or, with ``swapped=True``, of frame b: the Option-B
        call ``make_3d_2d_collection(stereogeom, uv_d, b_3dImage, uv, ...)`` of Cerebro.cpp:1562-1565, whose world points
        come from frame b.  Returns per pair (feature_position_uv [m,2] of frame a, feature_position_uv_d [m,2] of frame b,
        world_point [m,3])."""
        return self._collect(2, K, None, img3d) if swapped else self._collect(0, K, img3d, None)

    def make_3d_3d_collection(self, img3d_a: np.ndarray, img3d_b: np.ndarray):
        """make_3d_3d_collection__using__pfmatches_and_disparity: per pair (uv_X [m,3], uvd_Y [m,3])."""
        return self._collect(1, None, img3d_a, img3d_b)

    def _collect(self, mode, K, img_a, img_b):
        assert self._off1 is not None, "call match_gms first"
        n = len(self._off1) - 1
        if img_a is not None:
            img_a = np.ascontiguousarray(img_a, dtype=np.float32)
            assert img_a.ndim == 4 and img_a.shape[0] == n and img_a.shape[3] == 3
        if img_b is not None:
            img_b = np.ascontiguousarray(img_b, dtype=np.float32)
            assert img_b.ndim == 4 and img_b.shape[0] == n and img_b.shape[3] == 3
            assert img_a is None or img_b.shape == img_a.shape
        rows, cols = (img_a if img_a is not None else img_b).shape[1:3]
        tot = max(int(self._off1[-1]), 1)
        counts = np.zeros(n, dtype=np.int32)
        X = np.zeros((tot, 3))
        uv = np.zeros((tot, 2))
        uvd = np.zeros((tot, 2))
        Y = np.zeros((tot, 3))
        Kinv = np.ascontiguousarray(np.linalg.inv(np.asarray(K, dtype=np.float64))) if K is not None else None
        check(self._lib.cb_frontend_collect(self._h, mode, ptr(img_a), ptr(img_b), rows, cols, ptr(Kinv), ptr(counts), ptr(X), ptr(uv),
                                            ptr(uvd), ptr(Y)))
        out = []
        for p in range(n):
            a = int(self._off1[p])
            m = int(counts[p])
            if mode != 1:
                out.append((uv[a : a + m].copy(), uvd[a : a + m].copy(), X[a : a + m].copy()))
            else:
                out.append((X[a : a + m].copy(), Y[a : a + m].copy()))
        return out


class StaticPointFeatureMatching:
    """Name-for-name entry points of the reference class, single image pair, images in."""

    _fe = None

    @classmethod
    def _frontend(cls, n_feat):
        if cls._fe is None or cls._fe.max_features < n_feat:
            cls._fe = FrontEnd(max_pairs=1, max_features=max(n_feat, 5000))
        return cls._fe

    _feat = None

    @classmethod
    def _features(cls, rows, cols):
        from .features import Features

        if cls._feat is None or (cls._feat.rows, cls._feat.cols) != (rows, cols):
            cls._feat = Features(rows, cols, max_images=2, max_keypoints=8192)
        return cls._feat

    @classmethod
    def gms_point_feature_matches(cls, imleft_undistorted: np.ndarray, imright_undistorted: np.ndarray, n_orb_feat: int = 5000):
        """PointFeatureMatching.cpp:5-75: returns (u, ud), 3xN homogeneous pixel coordinates of the GMS matches (empty
        arrays when there is none).  ORB = ``cv::ORB::create(n_orb_feat)`` with FAST threshold 0 (:17-18) -- on the device
        (``cb_features_orb``), like the matcher and the GMS filter behind it."""
        a, b = np.asarray(imleft_undistorted), np.asarray(imright_undistorted)
        assert a.shape == b.shape and a.ndim == 2 and a.dtype == np.uint8
        r = cls._features(*a.shape).orb(np.stack([a, b]), n_orb_feat)
        if len(r[0]["pt"]) == 0 or len(r[1]["pt"]) == 0:
            return np.zeros((3, 0)), np.zeros((3, 0))
        fe = cls._frontend(max(len(r[0]["pt"]), len(r[1]["pt"])))
        h, w = a.shape
        fe.match_gms([r[0]["pt"]], [r[0]["desc"]], [r[1]["pt"]], [r[1]["desc"]], (w, h), (w, h))
        return fe.matched_points(0)
