"""Descriptor-only micro-benchmark (device-resident images, CUDA events)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_net  # noqa: E402
from cerebro_b200.descriptor import NetvladDescriptor  # noqa: E402


def main(batch=64, iters=5):
    net, raw, name = load_net()
    nd = NetvladDescriptor(net, 480, 640, 3, max_batch=batch)
    g = torch.Generator(device="cuda").manual_seed(1)
    imgs = torch.randint(0, 256, (batch, 480, 640, 3), generator=g, device="cuda", dtype=torch.uint8)
    out = nd.compute_device(imgs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        nd.compute_device(imgs, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print(json.dumps({"batch": batch, "ms": ms, "frames_per_s": batch / ms * 1e3, "gflops": 3.478 * batch / ms}))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
