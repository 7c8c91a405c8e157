"""ORACLE (test infrastructure, not product): CPU restatement of the reference's 3D-3D verifier
StaticTheiaPoseCompute::P3P_ICP = theia::Ransac<AlignPointCloudsUmeyamaWithRansac> (Option C of
process_loop_candidate_imagepair_consistent_pose_compute, src/Cerebro.cpp:1624-1629) and of the
three-way consistency check that turns the poses into a LoopEdge.

PARITY UNPINNED (Theia-SfM is absent, see oracle/dls_pnp.py).  Restated from the published algorithm
(S. Umeyama, "Least-squares estimation of transformation parameters between two point patterns",
PAMI 1991) and anchored on the reference call sites:
  * estimator                 src/DlsPnpWithRansac.h:117-166
      SampleSize() = 10 (:120); model accepted iff min(s, 1/s) > 0.9 (:139); the model keeps R and t of
      the similarity and DROPS the scale (:141-143); Error() = || R a_X + t - b_X || , f = 1 (:152-164:
      the `< 1 && > 8` condition can never hold)
  * wrapper                   src/DlsPnpWithRansac.cpp:16-122
      < 20 points -> -1 (:18-21); RansacParameters{error_thresh 0.1, min_inlier_ratio 0.7,
      max_iterations 50, min_iterations 5, use_mle true} (:88-93); returns summary.confidence (:120)
  * consistency + LoopEdge    src/ProcessedLoopCandidate.cpp:40-125, utils/PoseManipUtils.cpp:148-163

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import math

import numpy as np

from .dls_pnp import RansacParameters, _max_iterations, mle_cost, sample_indices


def umeyama(a: np.ndarray, b: np.ndarray):
    """theia::AlignPointCloudsUmeyama(left=a, right=b): b ~ s R a + t.  Returns (R, t, s)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = a.shape[0]
    mu_a, mu_b = a.mean(0), b.mean(0)
    da, db = a - mu_a, b - mu_b
    sigma = db.T @ da / n
    U, D, Vt = np.linalg.svd(sigma)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1.0
    R = U @ S @ Vt
    var_a = (da * da).sum() / n
    s = float((D * np.diag(S)).sum() / var_a)
    t = mu_b - s * R @ mu_a
    return R, t, s


def hypothesis(a, b, idx):
    """AlignPointCloudsUmeyamaWithRansac::EstimateModel (DlsPnpWithRansac.h:123-149)."""
    R, t, s = umeyama(a[idx], b[idx])
    if not (np.isfinite(s) and s > 0):
        return None
    if min(s, 1.0 / s) > 0.9:
        return R, t
    return None


def residuals(a, b, R, t):
    return np.linalg.norm(a @ R.T + t - b, axis=1)


def sample_table(seed: int, cand: int, n_hyp: int, n: int) -> np.ndarray:
    return np.stack([sample_indices(seed, cand, h, n, 10) for h in range(n_hyp)])


def ransac_icp(a: np.ndarray, b: np.ndarray, samples: np.ndarray, p: RansacParameters | None = None):
    """StaticTheiaPoseCompute::P3P_ICP with an explicit sample table [n_hyp, 10]."""
    p = p or RansacParameters(error_thresh=0.1, sample_size=10)
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    n = a.shape[0]
    T = np.eye(4)
    if n < 20:  # DlsPnpWithRansac.cpp:18-21
        return dict(confidence=-1.0, T=T, num_iterations=0, n_inliers=0, best_hyp=-1)
    log_fail = math.log(p.failure_probability)
    max_it = p.max_iterations
    if p.adaptive and p.min_inlier_ratio > 0:
        max_it = min(_max_iterations(p.sample_size, p.min_inlier_ratio, log_fail, p), p.max_iterations)
    best, best_cost, best_hyp, it = None, math.inf, -1, 0
    while it < max_it:
        model = hypothesis(a, b, samples[it]) if it < samples.shape[0] else None
        if model is not None:
            cost, inl = mle_cost(residuals(a, b, *model), p.error_thresh)
            if cost < best_cost:
                best, best_cost, best_hyp = model, cost, it
                ratio = inl.sum() / n
                if p.adaptive and ratio >= p.sample_size / n:
                    max_it = min(_max_iterations(p.sample_size, ratio, log_fail, p), max_it)
        it += 1
    if best is None:
        return dict(confidence=0.0, T=T, num_iterations=it, n_inliers=0, best_hyp=-1)
    _, inl = mle_cost(residuals(a, b, *best), p.error_thresh)
    ratio = inl.sum() / n
    T[:3, :3], T[:3, 3] = best
    return dict(confidence=float(1.0 - (1.0 - ratio**p.sample_size) ** it), T=T, num_iterations=it, n_inliers=int(inl.sum()), best_hyp=best_hyp)


# --------------------------------------------------------------------------------------------
# three-way consistency (ProcessedLoopCandidate::makeLoopEdgeMsgWithConsistencyCheck)
# --------------------------------------------------------------------------------------------
def R2ypr_deg(R: np.ndarray) -> np.ndarray:
    """PoseManipUtils::R2ypr (utils/PoseManipUtils.cpp:148-163): degrees."""
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = math.atan2(n[1], n[0])
    p = math.atan2(-n[2], n[0] * math.cos(y) + n[1] * math.sin(y))
    r = math.atan2(a[0] * math.sin(y) - a[1] * math.cos(y), -o[0] * math.sin(y) + o[1] * math.cos(y))
    return np.array([y, p, r]) / math.pi * 180.0


def consistency_check(op1, op2, icp, goodness, dt_sec: float, pf_matches: int):
    """Returns (publish: bool, pose_1T0 4x4 | None, weight | None), ProcessedLoopCandidate.cpp:40-125.
    Note the reference tests op1-icp twice and never uses the op1-op2 translation (:83-86)."""
    if abs(int(dt_sec)) < 10:  # :49-56 (ros::Duration::sec is the integer part)
        return False, None, None
    d12 = np.linalg.inv(op1) @ op2
    d1i = np.linalg.inv(op1) @ icp
    d2i = np.linalg.inv(op2) @ icp
    ypr_ok = all(np.abs(R2ypr_deg(d[:3, :3])).max() < 5.0 for d in (d12, d1i, d2i))  # :77-81
    tr_ok = np.abs(d1i[:3, 3]).max() < 0.2 and np.abs(d2i[:3, 3]).max() < 0.2  # :83-87
    if pf_matches > 800 and ypr_ok and tr_ok:  # :110
        return True, op1.copy(), float(max(goodness))  # :112-116
    return False, None, None


def synth_3d3d(rng: np.random.Generator, n=200, noise=0.01, outlier_frac=0.2, max_angle_deg=30.0, max_t=2.0):
    """3-D points in frame a and the same points in frame b (b = R a + t + noise), with gross outliers."""
    from .dls_pnp import ypr_to_R

    R = ypr_to_R(*np.deg2rad(rng.uniform(-max_angle_deg, max_angle_deg, 3)))
    t = rng.uniform(-max_t, max_t, 3)
    a = np.stack([rng.uniform(-6, 6, n), rng.uniform(-4, 4, n), rng.uniform(0.5, 20, n)], axis=1)
    b = a @ R.T + t + rng.normal(0, noise, (n, 3))
    nout = int(round(outlier_frac * n))
    if nout:
        bad = rng.choice(n, nout, replace=False)
        b[bad] += rng.uniform(-3, 3, (nout, 3))
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return a, b, T
