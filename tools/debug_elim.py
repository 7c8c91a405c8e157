"""Compares the action matrices of the two elimination kernels (CB_PNP_ELIM_V1=1 vs default) set by set."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200.pnp import PnpBatch  # noqa: E402
from oracle import dls_pnp as D  # noqa: E402

rng = np.random.default_rng(7)
n = 8
sets = [D.synth_candidate(rng, n=15, noise=1e-3)[:2] for _ in range(n)]
X = np.stack([s[0] for s in sets])
uv = np.stack([s[1] for s in sets])
S = {}
for v1 in ("1", "0"):
    os.environ["CB_PNP_ELIM_V1"] = v1
    pb = PnpBatch(max_candidates=1, max_points_total=n * 15, max_hypotheses=64)
    ns, R, t = pb.dls_minimal(X, uv)
    S[v1] = pb.debug_read(0, n)
    print("v1=" + v1, "n_solutions", ns)
    pb.close()
a, b = S["1"], S["0"]
print("identical:", np.array_equal(a, b))
for i in range(n):
    d = np.abs(a[i] - b[i])
    bad = np.argwhere(~(d == 0))
    rows = sorted(set(int(r) for r, _ in bad))
    print("set", i, "max |diff|", np.nanmax(d) if d.size else 0, "nan in v2", int(np.isnan(b[i]).sum()), "rows differing", rows)
    if i == 0 and rows:
        r = rows[0]
        print(" v1 row", r, a[i][r][:8])
        print(" v2 row", r, b[i][r][:8])
