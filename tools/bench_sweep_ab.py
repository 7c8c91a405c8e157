"""A/B of the tensor-core search sweep variants (device-resident inputs): sweep-kernel time from the library's own CUDA
events (cb_index_set_timing) and whole-search time from events around search_device.
  default = scores_tc2_kernel (tiled planes), CB_TC_V1=1 = scores_tc_kernel (row-major planes)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200.index import IndexFlatIP  # noqa: E402

PEAK = 6530.3
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]


def run(n, d, nq, env, iters=20):
    for k in ("CB_TC_V1", "CB_NO_TC", "CB_TC_Q64"):
        os.environ.pop(k, None)
    os.environ.update(env)
    g = torch.Generator(device="cuda").manual_seed(1)
    ix = IndexFlatIP(d, capacity=n)
    for a in range(0, n, 12_500):
        x = torch.randn((min(12_500, n - a), d), generator=g, device="cuda")
        x /= x.norm(dim=1, keepdim=True)
        ix.add(x)
    xq = torch.randn((nq, d), generator=g, device="cuda")
    xq /= xq.norm(dim=1, keepdim=True)
    out = ix.search_device(xq, 5)
    for _ in range(3):
        ix.search_device(xq, 5, out=out)
    torch.cuda.synchronize()
    ix.set_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ix.search_device(xq, 5, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms, cnt = ix.sweep_timing()
    ix.set_timing(False)
    sweep = ms / max(cnt, 1)
    gbs = n * d * 4 / sweep / 1e6
    print(json.dumps({"variant": env or "tiled (default)", "n": n, "d": d, "nq": nq, "sweep_ms": round(sweep, 4), "sweeps_per_search": cnt / iters,
                      "GBps": round(gbs, 1), "frac_hbm": round(gbs / PEAK, 3), "search_ms": round(e0.elapsed_time(e1) / iters, 4)}), flush=True)
    ix.close()


if __name__ == "__main__":
    for n, d in ((100_000, 8192), (12_500, 8192), (10_000, 4096)):
        for nq in (128, 64, 16):
            for env in ({}, {"CB_TC_Q64": "1"}):
                run(n, d, nq, env)
