"""Seeded synthetic inputs shared by tests and bench."""
import numpy as np


def unit_rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def planted_queries(db, targets, seed, score=0.9):
    """query_i = normalise(db[targets_i] * a + noise * b) with <q, db[target]> ~= score."""
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
