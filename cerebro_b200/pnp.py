"""Host-side mirror of the reference's geometric verifier.

``StaticTheiaPoseCompute.PNP`` keeps the name, argument order and return convention of
``StaticTheiaPoseCompute::PNP(w_X, c_uv_normalized, c_T_w, pnp__msg) -> float``
(src/DlsPnpWithRansac.h:169-179, src/DlsPnpWithRansac.cpp:132-245): the return value is the
RANSAC confidence, -1 when the input is refused (< 20 points), the pose comes back by reference.
``PnpBatch`` is the batched form the GPU wants: many loop candidates per call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import RansacParams, check, ptr


def default_params(**kw) -> RansacParams:
    p = RansacParams()
    _lib.load().cb_ransac_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class PnpBatch:
    def __init__(self, max_candidates: int = 64, max_points_total: int = 64 * 5000, max_hypotheses: int = 50, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.max_candidates = max_candidates
        check(self._lib.cb_pnp_create(C.byref(self._h), max_candidates, max_points_total, max_hypotheses, device))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.cb_pnp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, X_list, uv_list, params: RansacParams | None = None, samples=None):
        """X_list[c]: [n_c,3] float64, uv_list[c]: [n_c,2] float64 (normalised image coords).
        samples: optional int32 [n_cand, max_iterations, 15].
        Returns dict(T [n,4,4], confidence [n], num_iterations, n_inliers, best_hyp)."""
        params = params or default_params()
        n_cand = len(X_list)
        offsets = np.zeros(n_cand + 1, dtype=np.int32)
        offsets[1:] = np.cumsum([len(x) for x in X_list])
        X = np.ascontiguousarray(np.concatenate(X_list, axis=0), dtype=np.float64).reshape(-1, 3)
        uv = np.ascontiguousarray(np.concatenate(uv_list, axis=0), dtype=np.float64).reshape(-1, 2)
        if samples is not None:
            samples = np.ascontiguousarray(samples, dtype=np.int32)
            assert samples.shape == (n_cand, params.max_iterations, 15)
        T = np.empty((n_cand, 4, 4), dtype=np.float64)
        conf = np.empty(n_cand, dtype=np.float32)
        nit = np.empty(n_cand, dtype=np.int32)
        ninl = np.empty(n_cand, dtype=np.int32)
        bh = np.empty(n_cand, dtype=np.int32)
        check(
            self._lib.cb_pnp_solve_batch(
                self._h, n_cand, ptr(offsets), ptr(X), ptr(uv), C.byref(params), ptr(samples), ptr(T), ptr(conf),
                ptr(nit), ptr(ninl), ptr(bh),
            )
        )
        return dict(T=T, confidence=conf, num_iterations=nit, n_inliers=ninl, best_hyp=bh)

    def solve_device(self, offsets, X, uv, params: RansacParams, out=None, samples=None):
        """CUDA tensors: offsets int32 [n+1], X float64 [total,3], uv float64 [total,2].
        Asynchronous on torch's current stream."""
        import torch

        n_cand = offsets.numel() - 1
        dev = X.device
        if out is None:
            out = dict(
                T=torch.empty((n_cand, 4, 4), dtype=torch.float64, device=dev),
                confidence=torch.empty(n_cand, dtype=torch.float32, device=dev),
                num_iterations=torch.empty(n_cand, dtype=torch.int32, device=dev),
                n_inliers=torch.empty(n_cand, dtype=torch.int32, device=dev),
                best_hyp=torch.empty(n_cand, dtype=torch.int32, device=dev),
            )
        check(
            self._lib.cb_pnp_solve_batch_device(
                self._h, n_cand, ptr(offsets), X.shape[0], ptr(X), ptr(uv), C.byref(params), ptr(samples),
                ptr(out["T"]), ptr(out["confidence"]), ptr(out["num_iterations"]), ptr(out["n_inliers"]),
                ptr(out["best_hyp"]), _lib.current_stream_ptr(),
            )
        )
        return out

    def icp(self, A_list, B_list, params: RansacParams | None = None, samples=None):
        """Batched StaticTheiaPoseCompute::P3P_ICP: A_list[c], B_list[c] are [n_c,3] float64 (the same
        points in frames a and b).  Returns the same dict as ``solve`` (T = b_T_a)."""
        params = params or default_params(error_thresh=0.1)  # DlsPnpWithRansac.cpp:89
        n_cand = len(A_list)
        offsets = np.zeros(n_cand + 1, dtype=np.int32)
        offsets[1:] = np.cumsum([len(x) for x in A_list])
        A = np.ascontiguousarray(np.concatenate(A_list, axis=0), dtype=np.float64).reshape(-1, 3)
        B = np.ascontiguousarray(np.concatenate(B_list, axis=0), dtype=np.float64).reshape(-1, 3)
        if samples is not None:
            samples = np.ascontiguousarray(samples, dtype=np.int32)
            assert samples.shape == (n_cand, params.max_iterations, 10)
        T = np.empty((n_cand, 4, 4), dtype=np.float64)
        conf = np.empty(n_cand, dtype=np.float32)
        nit = np.empty(n_cand, dtype=np.int32)
        ninl = np.empty(n_cand, dtype=np.int32)
        bh = np.empty(n_cand, dtype=np.int32)
        check(
            self._lib.cb_pnp_icp_batch(
                self._h, n_cand, ptr(offsets), ptr(A), ptr(B), C.byref(params), ptr(samples), ptr(T), ptr(conf), ptr(nit),
                ptr(ninl), ptr(bh),
            )
        )
        return dict(T=T, confidence=conf, num_iterations=nit, n_inliers=ninl, best_hyp=bh)

    def dls_minimal(self, X_sets: np.ndarray, uv_sets: np.ndarray):
        """theia::DlsPnp on sets of exactly 15 points: X_sets [s,15,3], uv_sets [s,15,2].
        Returns (n_solutions [s], R [s,27,3,3], t [s,27,3])."""
        X_sets = np.ascontiguousarray(X_sets, dtype=np.float64)
        uv_sets = np.ascontiguousarray(uv_sets, dtype=np.float64)
        s, m = X_sets.shape[0], X_sets.shape[1]
        ns = np.empty(s, dtype=np.int32)
        R = np.zeros((s, 27, 3, 3), dtype=np.float64)
        t = np.zeros((s, 27, 3), dtype=np.float64)
        check(self._lib.cb_pnp_dls_minimal(self._h, s, m, ptr(X_sets), ptr(uv_sets), ptr(ns), ptr(R), ptr(t)))
        return ns, R, t

    def debug_read(self, what: int, n_sets: int) -> np.ndarray:
        """Intermediates of the last dls_minimal call: what = 0 the 27 x 27 action matrices, 1 the 60 gradient coefficients."""
        per = 729 if what == 0 else 60
        out = np.empty((n_sets, per), dtype=np.float64)
        n = self._lib.cb_pnp_debug_read(self._h, what, n_sets, ptr(out), out.size)
        if n < 0:
            check(int(n))
        return out.reshape(n_sets, 27, 27) if what == 0 else out


class StaticTheiaPoseCompute:
    """Same call shape as the reference class (src/DlsPnpWithRansac.h:169-179)."""

    _batch = None

    @classmethod
    def PNP(cls, w_X, c_uv_normalized, c_T_w: np.ndarray, pnp__msg: list | None = None, params=None, samples=None) -> float:
        """w_X: n x 3, c_uv_normalized: n x 2; c_T_w (4x4 array) is overwritten in place;
        pnp__msg (a list standing in for the by-reference string) gets the debug text appended.
        Returns the RANSAC confidence, or -1 when fewer than 20 points are given."""
        w_X = np.asarray(w_X, dtype=np.float64).reshape(-1, 3)
        c_uv = np.asarray(c_uv_normalized, dtype=np.float64).reshape(-1, 2)
        if w_X.shape[0] < 20:  # DlsPnpWithRansac.cpp:136-139 (no device work needed to refuse)
            return -1.0
        if cls._batch is None:
            cls._batch = PnpBatch(max_candidates=1, max_points_total=20000, max_hypotheses=4096)
        r = cls._batch.solve([w_X], [c_uv], params, None if samples is None else np.asarray(samples)[None])
        c_T_w[...] = r["T"][0]
        if pnp__msg is not None:
            pnp__msg.append(
                "DlsPnpWithRansac (best_rel_pose.b_T_a): %s;    num_iterations=%d  confidence=%f"
                % (np.array2string(r["T"][0], precision=6), int(r["num_iterations"][0]), float(r["confidence"][0]))
            )
        return float(r["confidence"][0])

    @classmethod
    def P3P_ICP(cls, uv_X, uvd_Y, uvd_T_uv: np.ndarray, p3p__msg: list | None = None, params=None, samples=None) -> float:
        """src/DlsPnpWithRansac.cpp:16-122: 3D-3D alignment (Umeyama) with RANSAC.  uv_X, uvd_Y: n x 3; the
        4x4 ``uvd_T_uv`` is overwritten; returns summary.confidence, or -1 for fewer than 20 points (:18-21)."""
        a = np.asarray(uv_X, dtype=np.float64).reshape(-1, 3)
        b = np.asarray(uvd_Y, dtype=np.float64).reshape(-1, 3)
        if a.shape[0] < 20:
            return -1.0
        if cls._batch is None:
            cls._batch = PnpBatch(max_candidates=1, max_points_total=20000, max_hypotheses=4096)
        r = cls._batch.icp([a], [b], params, None if samples is None else np.asarray(samples)[None])
        uvd_T_uv[...] = r["T"][0]
        if p3p__msg is not None:
            p3p__msg.append("ICP Ransac;     #iterations=%d    confidence=%f" % (int(r["num_iterations"][0]), float(r["confidence"][0])))
        return float(r["confidence"][0])
