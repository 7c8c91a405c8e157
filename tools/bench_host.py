"""Wall-clock of the individual host C-ABI calls (pinned inputs), to see what bounds the end-to-end path."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200 import synthetic  # noqa: E402
from cerebro_b200.descriptor import NetvladDescriptor  # noqa: E402
from cerebro_b200.index import IndexFlatIP  # noqa: E402
from cerebro_b200.keras_weights import random_mobilenet_netvlad  # noqa: E402
from cerebro_b200.pnp import PnpBatch, default_params  # noqa: E402


def timeit(fn, n=10):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    B = 64
    net = random_mobilenet_netvlad(seed=0)
    nd = NetvladDescriptor(net, 480, 640, 3, max_batch=B)
    imgs = torch.from_numpy(synthetic.band_limited_images(B, 480, 640, 3, seed=1)).pin_memory()
    a = imgs.numpy()
    print("desc.compute(64) ms        :", round(timeit(lambda: nd.compute(a)), 3))
    buf = torch.empty(imgs.shape, dtype=torch.uint8, device="cuda")
    def h2d():
        buf.copy_(imgs, non_blocking=True)
        torch.cuda.synchronize()
    ms = timeit(h2d)
    print("raw H2D 59 MB pinned ms    :", round(ms, 3), " GB/s", round(imgs.numel() / ms / 1e6, 1))
    d_dev = torch.empty((B, 8192), dtype=torch.float32, device="cuda")
    idev = imgs.cuda()
    print("desc.compute_device(64) ms :", round(timeit(lambda: (nd.compute_device(idev, out=d_dev), torch.cuda.synchronize())), 3))
    n, d = 100_000, 8192
    ix = IndexFlatIP(d, capacity=n)
    g = torch.Generator(device="cuda").manual_seed(1)
    for c in range(0, n, 12500):
        x = torch.randn((12500, d), generator=g, device="cuda")
        x /= x.norm(dim=1, keepdim=True)
        ix.add_local(x)
    q = torch.randn((B, d)).numpy().astype(np.float32)
    print("index.search(64, host) ms  :", round(timeit(lambda: ix.search(q, 5)), 3))
    rng = np.random.default_rng(5)
    cands = [synthetic.loop_candidate(rng, n=200) for _ in range(B)]
    Xs = [c[0] for c in cands]
    uvs = [c[1] for c in cands]
    pb = PnpBatch(B, B * 200, 50)
    prm = default_params()
    print("pnp.solve(64, host) ms     :", round(timeit(lambda: pb.solve(Xs, uvs, prm)), 3))


if __name__ == "__main__":
    main()
