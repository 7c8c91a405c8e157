"""Bring-up aid: per-layer comparison of the TMA-halo fused kernel against the global-memory fused kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import golden_io  # noqa: E402
import synth  # noqa: E402
from cerebro_b200.descriptor import NetvladDescriptor  # noqa: E402
from cerebro_b200.keras_weights import fold_mobilenet_netvlad  # noqa: E402


def act(net, imgs, h, w, c, layer, no_halo):
    os.environ["CB_DEBUG_STOP_LAYER"] = str(layer)
    if no_halo:
        os.environ["CB_NO_HALO"] = "1"
    else:
        os.environ.pop("CB_NO_HALO", None)
    nd = NetvladDescriptor(net, h, w, c, max_batch=imgs.shape[0])
    nd.compute(imgs)
    a = nd.get_activation(layer)
    nd.close()
    return a


def main():
    h, w, c = (int(sys.argv[1]), int(sys.argv[2]), 3) if len(sys.argv) > 2 else (96, 128, 3)
    raw = golden_io.raw_weights("mobilenet_conv7")
    net = fold_mobilenet_netvlad(raw)
    imgs = synth.band_limited_images(1, h, w, c, seed=5)
    H, W = h // 2, w // 2
    chans = [64, 128, 128, 256, 256, 512, 512]
    strides = [1, 2, 1, 2, 1, 2, 1]
    for bi in range(7):
        layer = 2 * (bi + 1)
        if strides[bi] == 2:
            H, W = H // 2, W // 2
        C = chans[bi]
        a = act(net, imgs, h, w, c, layer, False)
        b = act(net, imgs, h, w, c, layer, True)
        d = np.abs(a - b).reshape(H, W, C)
        print("block %d layer %d  %dx%dx%d  max err %.4g  frac>0.05 %.4f" % (bi + 1, layer, H, W, C, d.max(), (d > 0.05).mean()))
        if d.max() > 0.05:
            bad = d > 0.05
            print("  bad by y%16:", np.round([bad[y::16].mean() for y in range(min(16, H))], 2))
            print("  bad by x%8 :", np.round([bad[:, x::8].mean() for x in range(min(8, W))], 2))
            print("  bad by ch/8 (first 16):", np.round([bad[:, :, k * 8:(k + 1) * 8].mean() for k in range(min(16, C // 8))], 2))
            print("  sample a:", a.reshape(H, W, C)[1, 1, :8], "\n  sample b:", b.reshape(H, W, C)[1, 1, :8])
            break


if __name__ == "__main__":
    main()
