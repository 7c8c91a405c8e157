"""Seeded synthetic inputs for benches and tests: descriptor rows,
planted queries, band-limited images, 3D-2D loop-candidate correspondences.  numpy only."""
from __future__ import annotations

import math

import numpy as np


def unit_rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def planted_queries(db, targets, seed, score=0.9):
    """query_i = normalise(score * db[targets_i] + sqrt(1-score^2) * noise): <q, db[target]> ~= score."""
    rng = np.random.default_rng(seed)
    d = db.shape[1]
    noise = rng.standard_normal((len(targets), d)).astype(np.float32)
    noise /= np.linalg.norm(noise, axis=1, keepdims=True)
    q = score * db[targets] + np.sqrt(1 - score**2) * noise
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def band_limited_images(n, h, w, c, seed):
    """uint8 images with spatial structure (smoothed noise), deterministic."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, h, w, c), dtype=np.uint8)
    for i in range(n):
        lo = rng.standard_normal((h // 8 + 2, w // 8 + 2, c))
        up = np.kron(lo, np.ones((8, 8, 1)))[:h, :w]
        hi = rng.standard_normal((h, w, c)) * 0.35
        img = up + hi
        img = (img - img.min()) / (img.max() - img.min())
        out[i] = (img * 255).astype(np.uint8)
    return out


def ypr_to_R(y, p, r):
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    return Rz @ Ry @ Rx


def loop_candidate(rng, n=200, noise=1e-3, outlier_frac=0.2, max_angle_deg=30.0, max_t=2.0):
    """3-D points seen by camera b (depth 0.5-20 m, inside the reference's 0.1-25 m gate,
    utils/PointFeatureMatching.cpp:125) expressed in frame a through a random pose.
    Returns (X_a [n,3], uv_b [n,2] normalised image coords, T_b_a [4,4], inlier mask)."""
    ang = np.deg2rad(rng.uniform(-max_angle_deg, max_angle_deg, 3))
    R = ypr_to_R(*ang)
    t = rng.uniform(-max_t, max_t, 3)
    z = rng.uniform(0.5, 20.0, n)
    uvt = np.stack([rng.uniform(-0.6, 0.6, n), rng.uniform(-0.45, 0.45, n)], axis=1)
    Pb = np.concatenate([uvt * z[:, None], z[:, None]], axis=1)
    Xa = (Pb - t) @ R
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


def textured_scenes(n, h, w, c, seed):
    """uint8 images whose LOCAL statistics differ from image to image (so that whole-image descriptors differ, which
    band-limited noise does not achieve): a few regions (nearest-seed cells), each filled with an oriented grating /
    checkerboard / block noise / ramp of random scale, orientation, contrast and colour, plus sensor-like noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty((n, h, w, c), dtype=np.uint8)
    for i in range(n):
        k = int(rng.integers(3, 7))
        cy, cx = rng.uniform(0, h, k), rng.uniform(0, w, k)
        cell = np.argmin((yy[None] - cy[:, None, None]) ** 2 + (xx[None] - cx[:, None, None]) ** 2, axis=0)
        img = np.zeros((h, w, c), dtype=np.float32)
        for r in range(k):
            kind = int(rng.integers(0, 4))
            scale = float(rng.uniform(3.0, 40.0))
            th = float(rng.uniform(0, np.pi))
            u = (xx * np.cos(th) + yy * np.sin(th)) / scale
            v = (-xx * np.sin(th) + yy * np.cos(th)) / scale
            if kind == 0:
                t = 0.5 + 0.5 * np.sin(2 * np.pi * u + rng.uniform(0, 6.28))
            elif kind == 1:
                t = ((np.floor(u) + np.floor(v)) % 2).astype(np.float32)
            elif kind == 2:
                lo = rng.standard_normal((int(h / scale) + 3, int(w / scale) + 3)).astype(np.float32)
                iy = np.clip((yy / scale).astype(int), 0, lo.shape[0] - 1)
                ix = np.clip((xx / scale).astype(int), 0, lo.shape[1] - 1)
                t = (lo[iy, ix] > 0).astype(np.float32)
            else:
                t = (u - u.min()) / (u.max() - u.min() + 1e-6)
            base = rng.uniform(20, 200, c).astype(np.float32)
            amp = rng.uniform(20, 120) * rng.uniform(0.3, 1.0, c).astype(np.float32)
            tex = base[None, None, :] + amp[None, None, :] * t[:, :, None]
            m = cell == r
            img[m] = tex[m]
        img += rng.standard_normal((h, w, c)).astype(np.float32) * 4.0
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out
