"""Micro-benchmark of the search sweep (device-resident inputs, CUDA events)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200.index import IndexFlatIP  # noqa: E402

PEAK = 6566.7
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def run(n, d, nq, iters=10):
    g = torch.Generator(device="cuda").manual_seed(1)
    db = torch.randn((n, d), generator=g, device="cuda")
    db /= db.norm(dim=1, keepdim=True)
    xq = db[torch.arange(nq, device="cuda") * (n // nq)].contiguous()
    ix = IndexFlatIP(d, capacity=n)
    ix.add(db)
    del db
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = ix.search_device(xq, 5)
    for _ in range(3):
        ix.search_device(xq, 5, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.search_device(xq, 5, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    sweeps = (nq + 15) // 16
    gbs = n * d * 4 * sweeps / ms / 1e6
    print(json.dumps({"n": n, "d": d, "nq": nq, "ms": round(ms, 4), "sweeps": sweeps, "GBps_per_sweep": round(gbs, 1),
                      "frac_hbm": round(gbs / PEAK, 3), "queries_per_s": round(nq / ms * 1e3, 1)}))
    ix.close()


if __name__ == "__main__":
    for n, d in ((10_000, 4096), (100_000, 8192)):
        for nq in (1, 3, 8, 16, 64):
            run(n, d, nq)
