"""ORACLE (test infrastructure, not product): CPU restatement of the reference's
geometric verifier  StaticTheiaPoseCompute::PNP  =  theia::Ransac<DlsPnpWithRansac>.

PARITY UNPINNED: the arithmetic lives in Theia-SfM, an un-vendored, un-pinned
third-party dependency (``find_package(Theia REQUIRED)`` CMakeLists.txt:27,
``#include <theia/theia.h>`` src/DlsPnpWithRansac.h:28) whose source is absent
from /root/reference, and the reference holds no golden vectors for this path.
The solver below restates the PUBLISHED algorithm -- J. Hesch & S. Roumeliotis,
"A Direct Least-Squares (DLS) Method for PnP", ICCV 2011 -- and Theia's
documented RANSAC bookkeeping, and is pinned by (i) noise-free known-pose
recovery to 1e-9, (ii) agreement with cv2.solvePnP (SQPNP / iterative) on noisy
data, (iii) the Bezout count of 27 roots with residual checks on every root.

What is anchored on reference call sites:
  * estimator contract            src/DlsPnpWithRansac.h:42-100
      SampleSize() = 15 (:45); a hypothesis yields a model only when DlsPnp returns
      EXACTLY ONE solution (:62-71); Error() = |x/z-u| + |y/z-v|, f = 1 (:75-99)
  * wrapper                       src/DlsPnpWithRansac.cpp:132-245
      < 20 points -> return -1 (:136-139); RansacParameters{error_thresh 0.03,
      min_inlier_ratio 0.7, max_iterations 50, min_iterations 5, use_mle true}
      (:207-212); returns summary.confidence (:240), pose by reference (:239)

DLS-PnP (paper sections 3-4), as implemented here:
  model   alpha_i f_i = C r_i + t ,  f_i = normalise([u_i, v_i, 1])
  1. H = (n I - sum f f^T)^-1 ;  t = T vec(C),  T = H sum (f f^T - I) L(r_i),  L(r) vec(C) = C r
  2. J = vec(C)^T G vec(C),  G = sum (L_i+T)^T (I - f f^T) (L_i+T)
  3. Cayley:  (1+s.s) C = (1-s.s) I + 2[s]x + 2 s s^T  =: Cbar(s), quadratic in s;
     J' = (1+s.s)^2 J = m(s)^T Q m(s) with m the 10 monomials of degree <= 2
  4. grad J' = 0 : three cubics f1,f2,f3 in (s1,s2,s3), 27 roots (Bezout)
  5. Macaulay resultant with a generic linear form f0: 120x120 matrix over the
     monomials of degree <= 7; Schur complement on the 93 non-reduced monomials gives
     the 27x27 matrix of multiplication by f0 in the quotient ring; its eigenvectors
     are the reduced monomials evaluated at the roots -> (s1,s2,s3)
  6. keep real roots; rebuild C, t; drop solutions with any point behind the camera.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------
# polynomial index tables (built once)
# --------------------------------------------------------------------------
# the 10 monomials of degree <= 2 that Cbar is linear in
M2 = [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
# cubic monomials (degree <= 3), 20 of them, fixed order
M3 = [e for e in itertools.product(range(4), repeat=3) if sum(e) <= 3]
M3_IDX = {e: i for i, e in enumerate(M3)}
# all monomials of degree <= 7 : 120
M7 = [e for e in itertools.product(range(8), repeat=3) if sum(e) <= 7]


def _is_reduced(e):
    return max(e) <= 2


# basis of the quotient ring: s1^a s2^b s3^c, a,b,c <= 2, index 9a+3b+c
BASIS = sorted([e for e in M7 if _is_reduced(e)], key=lambda e: 9 * e[0] + 3 * e[1] + e[2])
# the 93 non-reduced monomials, ordered by DESCENDING total degree: with this order the
# 93x93 block is block upper-triangular (rows of degree p touch only columns of degree <= p)
NONRED = sorted([e for e in M7 if not _is_reduced(e)], key=lambda e: (-sum(e), e))
COLS = BASIS + NONRED
COL_IDX = {e: i for i, e in enumerate(COLS)}
assert len(BASIS) == 27 and len(NONRED) == 93 and len(M3) == 20

# generic linear form f0 = U0 + U1 s1 + U2 s2 + U3 s3 (Theia draws it at random; the
# roots do not depend on it).  Fixed here so oracle and kernel are reproducible.
F0 = np.array([0.3721, 1.1193, -0.8647, 0.6158])

# A: vec(Cbar) (column-major, like Eigen's .data()) = A @ m(s)
_A = np.zeros((9, 10))


def _set(r, c, terms):
    for mono, val in terms:
        _A[c * 3 + r, M2.index(mono)] += val


_s11, _s12, _s13, _s22, _s23, _s33 = (2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2)
_s1, _s2, _s3, _one = (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)
_set(0, 0, [(_one, 1), (_s11, 1), (_s22, -1), (_s33, -1)])
_set(0, 1, [(_s12, 2), (_s3, -2)])
_set(0, 2, [(_s13, 2), (_s2, 2)])
_set(1, 0, [(_s12, 2), (_s3, 2)])
_set(1, 1, [(_one, 1), (_s11, -1), (_s22, 1), (_s33, -1)])
_set(1, 2, [(_s23, 2), (_s1, -2)])
_set(2, 0, [(_s13, 2), (_s2, -2)])
_set(2, 1, [(_s23, 2), (_s1, 2)])
_set(2, 2, [(_one, 1), (_s11, -1), (_s22, -1), (_s33, 1)])
A_CAYLEY = _A.copy()


def _build_grad_table():
    """E[k, mu, a, b]:  d/ds_k (m^T Q m) = sum E[k,mu,a,b] Q[a,b] * monomial mu."""
    E = np.zeros((3, 20, 10, 10))
    for k in range(3):
        for a, ea in enumerate(M2):
            if ea[k] == 0:
                continue
            c = ea[k]
            da = list(ea)
            da[k] -= 1
            for b, eb in enumerate(M2):
                mu = tuple(x + y for x, y in zip(da, eb))
                # d(m_a m_b) = dm_a m_b + m_a dm_b ; by symmetry of Q the two halves are equal
                E[k, M3_IDX[mu], a, b] += 2.0 * c
    return E.reshape(3, 20, 100)


GRAD_TABLE = _build_grad_table()


def _build_macaulay_tables():
    """Index tables: M[row, col] += coeff for every (row, col, source) triple."""
    rows, cols, src = [], [], []  # src: 0..59 -> f_k coeff (k*20+mu), 60..63 -> F0
    # basis rows: f0 * b
    f0_terms = [((0, 0, 0), 60), ((1, 0, 0), 61), ((0, 1, 0), 62), ((0, 0, 1), 63)]
    for r, b in enumerate(BASIS):
        for e, s in f0_terms:
            mono = tuple(x + y for x, y in zip(b, e))
            rows.append(r)
            cols.append(COL_IDX[mono])
            src.append(s)
    # non-reduced rows: monomial x^alpha, first i with alpha_i >= 3 -> f_i * x^(alpha - 3 e_i)
    for j, al in enumerate(NONRED):
        i = next(i for i in range(3) if al[i] >= 3)
        mult = list(al)
        mult[i] -= 3
        for mu, e in enumerate(M3):
            mono = tuple(x + y for x, y in zip(mult, e))
            rows.append(27 + j)
            cols.append(COL_IDX[mono])
            src.append(i * 20 + mu)
    return np.array(rows), np.array(cols), np.array(src)


MAC_ROWS, MAC_COLS, MAC_SRC = _build_macaulay_tables()
IDX_S1, IDX_S2, IDX_S3, IDX_ONE = 9, 3, 1, 0  # positions of s1, s2, s3, 1 in BASIS


# --------------------------------------------------------------------------
# solver
# --------------------------------------------------------------------------
def cayley_to_rotation(s):
    s = np.asarray(s, dtype=np.float64)
    ss = float(s @ s)
    sx = np.array([[0, -s[2], s[1]], [s[2], 0, -s[0]], [-s[1], s[0], 0]])
    return ((1 - ss) * np.eye(3) + 2 * sx + 2 * np.outer(s, s)) / (1 + ss)


def dls_setup(X: np.ndarray, uv: np.ndarray):
    """Steps 1-4: returns (T [3,9], grad coefficients [3,20], Q [10,10])."""
    X = np.asarray(X, dtype=np.float64)
    uv = np.asarray(uv, dtype=np.float64)
    n = X.shape[0]
    f = np.concatenate([uv, np.ones((n, 1))], axis=1)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    F = f[:, :, None] * f[:, None, :]  # n x 3 x 3
    I3 = np.eye(3)
    H = np.linalg.inv(n * I3 - F.sum(0))
    # L_i = [r1 I, r2 I, r3 I]
    L = np.concatenate([X[:, 0, None, None] * I3, X[:, 1, None, None] * I3, X[:, 2, None, None] * I3], axis=2)  # n x 3 x 9
    T = H @ np.einsum("nij,njk->ik", F - I3, L)
    LT = L + T
    G = np.einsum("nji,njk,nkl->il", LT, I3 - F, LT)  # 9 x 9
    Q = A_CAYLEY.T @ G @ A_CAYLEY
    Q = 0.5 * (Q + Q.T)
    coef = GRAD_TABLE @ Q.reshape(100)  # 3 x 20
    return T, coef, Q


def macaulay_matrix(coef: np.ndarray) -> np.ndarray:
    src = np.concatenate([coef.reshape(60), F0])
    M = np.zeros((120, 120))
    np.add.at(M, (MAC_ROWS, MAC_COLS), src[MAC_SRC])
    return M


def action_matrix(coef: np.ndarray) -> np.ndarray:
    """Schur complement: multiplication-by-f0 matrix in the 27-dim quotient ring."""
    M = macaulay_matrix(coef)
    return M[:27, :27] - M[:27, 27:] @ np.linalg.solve(M[27:, 27:], M[27:, :27])


def dls_roots(coef: np.ndarray):
    """All 27 complex roots (s1,s2,s3) of grad J' = 0 and the real-root mask."""
    S = action_matrix(coef)
    lam, V = np.linalg.eig(S)
    s = np.stack([V[IDX_S1] / V[IDX_ONE], V[IDX_S2] / V[IDX_ONE], V[IDX_S3] / V[IDX_ONE]], axis=1)
    # LAPACK returns exactly-real eigenvectors for the 1x1 blocks of the real Schur form;
    # complex-conjugate pairs have non-zero imaginary parts.  Real root <=> real eigenvalue.
    real = lam.imag == 0.0
    return s, real


def dls_pnp(X: np.ndarray, uv: np.ndarray):
    """theia::DlsPnp as called at DlsPnpWithRansac.h:61.  Returns list of (C [3,3], t [3])
    with  depth * [u,v,1] ~ C X + t,  every real stationary point of J' whose points all
    have non-negative depth."""
    T, coef, _ = dls_setup(X, uv)
    s, real = dls_roots(coef)
    sols = []
    for j in np.nonzero(real)[0]:
        sj = s[j].real
        if not np.all(np.isfinite(sj)):
            continue
        C = cayley_to_rotation(sj)
        t = T @ C.reshape(9, order="F")
        z = X @ C[2] + t[2]
        if np.all(z >= 0):
            sols.append((C, t))
    return sols


# --------------------------------------------------------------------------
# RANSAC (theia::Ransac with MLE quality measurement), SURVEY.md appendix A.2
# --------------------------------------------------------------------------
@dataclass
class RansacParameters:  # DlsPnpWithRansac.cpp:207-212 + Theia defaults
    error_thresh: float = 0.03
    min_inlier_ratio: float = 0.7
    max_iterations: int = 50
    min_iterations: int = 5
    use_mle: bool = True
    failure_probability: float = 0.01
    sample_size: int = 15  # DlsPnpWithRansac.h:45
    adaptive: bool = True  # False = evaluate exactly max_iterations hypotheses (BASELINE config 5)


_MASK = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15


def _mix(z: int) -> int:
    z &= _MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def sample_indices(seed: int, cand: int, hyp: int, n: int, k: int = 15) -> np.ndarray:
    """Counter-based sampler shared bit-for-bit with the CUDA kernel: k distinct indices
    in [0,n) for (seed, candidate, hypothesis); duplicates are redrawn."""
    key = _mix(seed + _GOLD * (cand + 1))
    key = _mix(key + _GOLD * (hyp + 1))
    out = []
    ctr = 0
    while len(out) < k:
        ctr += 1
        r = _mix(key + _GOLD * ctr)
        idx = ((r >> 32) * n) >> 32
        if idx not in out:
            out.append(idx)
    return np.array(out, dtype=np.int32)


def sample_table(seed: int, cand: int, n_hyp: int, n: int, k: int = 15) -> np.ndarray:
    return np.stack([sample_indices(seed, cand, h, n, k) for h in range(n_hyp)])


def residuals(X, uv, C, t):
    """DlsPnpWithRansac::Error (DlsPnpWithRansac.h:75-99), vectorised; f = 1."""
    P = X @ C.T + t
    return np.abs(P[:, 0] / P[:, 2] - uv[:, 0]) + np.abs(P[:, 1] / P[:, 2] - uv[:, 1])


def mle_cost(res, thresh):
    inl = res < thresh
    return float(np.where(inl, res, thresh).sum()), inl


def _max_iterations(sample_size, inlier_ratio, log_failure_prob, p: RansacParameters) -> int:
    if inlier_ratio == 1.0:
        return p.min_iterations
    log_prob = math.log(1.0 - inlier_ratio**sample_size) - np.finfo(np.float64).eps
    num = log_failure_prob / log_prob
    return int(max(float(p.min_iterations), min(num, float(p.max_iterations))))


def hypothesis(X, uv, idx):
    """EstimateModel (DlsPnpWithRansac.h:48-72): model iff exactly one DLS solution."""
    sols = dls_pnp(X[idx], uv[idx])
    if len(sols) == 1:
        return sols[0]
    return None


def ransac_pnp(X: np.ndarray, uv: np.ndarray, samples: np.ndarray, p: RansacParameters | None = None):
    """StaticTheiaPoseCompute::PNP (DlsPnpWithRansac.cpp:132-245) with an externally
    supplied sample table ``samples`` [n_hyp, 15] (row h = hypothesis h).

    Returns dict(confidence, T [4,4], num_iterations, n_inliers, best_hyp, best_cost)
    confidence = -1 when refused (< 20 points)."""
    p = p or RansacParameters()
    X = np.asarray(X, dtype=np.float64)
    uv = np.asarray(uv, dtype=np.float64)
    n = X.shape[0]
    Tm = np.eye(4)
    if n < 20:  # :136-139
        return dict(confidence=-1.0, T=Tm, num_iterations=0, n_inliers=0, best_hyp=-1, best_cost=math.inf)
    log_fail = math.log(p.failure_probability)
    max_it = p.max_iterations
    if p.adaptive and p.min_inlier_ratio > 0:
        max_it = min(_max_iterations(p.sample_size, p.min_inlier_ratio, log_fail, p), p.max_iterations)
    best_cost = math.inf
    best = None
    best_hyp = -1
    it = 0
    while it < max_it:
        model = hypothesis(X, uv, samples[it]) if it < samples.shape[0] else None
        if model is not None:
            res = residuals(X, uv, *model)
            cost, inl = mle_cost(res, p.error_thresh)
            if cost < best_cost:
                best, best_cost, best_hyp = model, cost, it
                ratio = inl.sum() / n
                if p.adaptive and ratio >= p.sample_size / n:
                    max_it = min(_max_iterations(p.sample_size, ratio, log_fail, p), max_it)
        it += 1
    num_iterations = it
    if best is None:
        return dict(confidence=0.0, T=Tm, num_iterations=num_iterations, n_inliers=0, best_hyp=-1, best_cost=math.inf)
    res = residuals(X, uv, *best)
    _, inl = mle_cost(res, p.error_thresh)
    ratio = inl.sum() / n
    conf = 1.0 - (1.0 - ratio**p.sample_size) ** num_iterations
    Tm[:3, :3] = best[0]
    Tm[:3, 3] = best[1]
    return dict(confidence=float(conf), T=Tm, num_iterations=num_iterations, n_inliers=int(inl.sum()), best_hyp=best_hyp, best_cost=best_cost)


# --------------------------------------------------------------------------
# synthetic data
# --------------------------------------------------------------------------
def ypr_to_R(y, p, r):
    """PoseManipUtils::ypr2R (utils/PoseManipUtils.cpp:165-191), radians here."""
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    return Rz @ Ry @ Rx


def synth_candidate(rng: np.random.Generator, n: int = 200, noise: float = 1e-3, outlier_frac: float = 0.2, max_angle_deg: float = 30.0, max_t: float = 2.0):
    """3D points in camera-b frustum (depth 0.5-20 m), expressed in frame a through a
    random pose; returns (X_a [n,3], uv_b [n,2], T_b_a [4,4], inlier mask)."""
    ang = np.deg2rad(rng.uniform(-max_angle_deg, max_angle_deg, 3))
    R = ypr_to_R(*ang)
    t = rng.uniform(-max_t, max_t, 3)
    z = rng.uniform(0.5, 20.0, n)
    uvt = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.45, 0.45, n)], axis=1)
    Pb = np.concatenate([uvt * z[:, None], z[:, None]], axis=1)
    Xa = (Pb - t) @ R  # R^T (Pb - t)
    uv = uvt + rng.normal(0, noise, (n, 2))
    nout = int(round(outlier_frac * n))
    mask = np.ones(n, dtype=bool)
    if nout:
        bad = rng.choice(n, nout, replace=False)
        uv[bad] = np.stack([rng.uniform(-0.6, 0.6, nout), rng.uniform(-0.45, 0.45, nout)], axis=1)
        mask[bad] = False
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return Xa, uv, T, mask


def pose_error(Ta, Tb):
    """(rotation angle [rad], translation distance [m]) between two 4x4 poses."""
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    c = max(-1.0, min(1.0, (np.trace(dR) - 1) / 2))
    return math.acos(c), float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))
