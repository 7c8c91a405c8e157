"""Descriptor-only micro-benchmark (device-resident images, CUDA events).
  python tools/bench_desc.py [batch]            total forward time
  python tools/bench_desc.py [batch] --layers   per-kernel times from cumulative runs stopped after every fused block
                                                (CB_DEBUG_STOP_LAYER), live and un-serialised, unlike an ncu launch list"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_net  # noqa: E402
from cerebro_b200.descriptor import NetvladDescriptor  # noqa: E402


def timed(nd, imgs, out, iters):
    nd.compute_device(imgs, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        nd.compute_device(imgs, out=out)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main(batch=64, iters=10, layers=False):
    net, raw, name = load_net()
    g = torch.Generator(device="cuda").manual_seed(1)
    imgs = torch.randint(0, 256, (batch, 480, 640, 3), generator=g, device="cuda", dtype=torch.uint8)
    out = torch.empty((batch, 8192), dtype=torch.float32, device="cuda")
    res = {"batch": batch}
    if layers:
        cum = {}
        stops = [0] + [2 * (i + 1) for i in range(len(net["blocks"]))]
        for s in stops:
            os.environ["CB_DEBUG_STOP_LAYER"] = str(s)
            nd = NetvladDescriptor(net, 480, 640, 3, max_batch=batch)
            cum[s] = timed(nd, imgs, out, iters)
            nd.close()
        os.environ.pop("CB_DEBUG_STOP_LAYER", None)
        prev = 0.0
        per = {}
        for s in stops:
            per["stem" if s == 0 else "block%d" % (s // 2)] = round((cum[s] - prev) * 1e3, 1)
            prev = cum[s]
        res["us_per_kernel"] = per
    nd = NetvladDescriptor(net, 480, 640, 3, max_batch=batch)
    ms = timed(nd, imgs, out, iters)
    if layers:
        res["us_per_kernel"]["vlad_head"] = round((ms - prev) * 1e3, 1)
    res.update({"ms": ms, "frames_per_s": batch / ms * 1e3, "gflops": 3.478 * batch / ms})
    print(json.dumps(res))


if __name__ == "__main__":
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    main(int(a[0]) if a else 64, layers="--layers" in sys.argv)
