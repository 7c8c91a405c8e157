"""Micro-benchmark of the batched DLS-PnP RANSAC (device-resident inputs, CUDA events)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200.pnp import PnpBatch, default_params  # noqa: E402
from oracle import dls_pnp as D  # noqa: E402


def run(n_cand, H, n_pts=200, iters=3):
    rng = np.random.default_rng(3)
    base = [D.synth_candidate(rng, n=n_pts) for _ in range(min(n_cand, 16))]
    Xs = [base[i % len(base)][0] for i in range(n_cand)]
    uvs = [base[i % len(base)][1] for i in range(n_cand)]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum([len(x) for x in Xs])]), dtype=torch.int32, device="cuda")
    X = torch.tensor(np.concatenate(Xs), device="cuda")
    uv = torch.tensor(np.concatenate(uvs), device="cuda")
    pb = PnpBatch(max_candidates=n_cand, max_points_total=n_cand * n_pts, max_hypotheses=H)
    prm = default_params(seed=1, max_iterations=H, adaptive=0)
    out = pb.solve_device(offsets, X, uv, prm)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pb.solve_device(offsets, X, uv, prm, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    hyps = n_cand * H
    print(json.dumps({"n_cand": n_cand, "H": H, "n_pts": n_pts, "ms": round(ms, 3), "hyp_per_s": round(hyps / ms * 1e3),
                      "cand_per_s": round(n_cand / ms * 1e3, 1), "conf_mean": float(out["confidence"].mean())}))
    pb.close()


if __name__ == "__main__":
    run(64, 50)
    run(1024, 50)
    run(64, 4096)
    if "--full" in sys.argv:
        run(1024, 4096)
