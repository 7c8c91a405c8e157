"""Generates tests/golden/* in the build container (where /root/reference is mounted).

  keras_raw_<model>.npz   RAW Keras weights of four shipped models, re-packed (not sources: data
                          the parity tests need on the GPU box, where /root/reference is absent)
  netvlad_golden.npz      seeded uint8 images -> descriptors from oracle/netvlad.py in fp64
  pnp_golden.npz          seeded candidates + sample tables -> oracle RANSAC outputs
The oracle, not the reference, produced the outputs (the reference cannot run here: Keras 2.2.4
/ TF 1.11 / Theia are not installable) -- see DESIGN.md "parity unpinned".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cerebro_b200.keras_weights import load_keras_file  # noqa: E402
from oracle import dls_pnp, netvlad  # noqa: E402
from tests import synth  # noqa: E402

REF = "/root/reference/scripts/keras.models/"
MODELS = {
    "mobilenet_conv7": REF + "mobilenet_conv7_allpairloss.keras",
    "gray_conv6": REF + "Apr2019/gray_conv6_K16__centeredinput/core_model.1000.keras",
    # June2019 models: MobileNetV2 prefix (inverted residuals, 1024-D) and MobileNet-v1 cut after conv_pw_6
    "mobilenetv2_block9_gray": REF + "June2019/centeredinput-m1to1-240x320x1__mobilenetv2-block_9_add__K16__allpairloss/modelarch_and_weights.2000.h5",
    "mobilenet_pw6": REF + "June2019/centeredinput-m1to1-240x320x3__mobilenet-conv_pw_6_relu__K16__allpairloss/modelarch_and_weights.700.h5",
}
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    gold = {}
    for name, path in MODELS.items():
        w = load_keras_file(path)
        np.savez(os.path.join(OUT, "keras_raw_%s.npz" % name), **{k.replace("/", "__"): v for k, v in w.items()})
        c = (w["conv1/kernel"] if "conv1/kernel" in w else w["Conv1/kernel"]).shape[2]
        sizes = [(96, 128), (240, 320)]
        # the benchmarked configuration (480x640, BASELINE configs 2/4) and EuRoC's 480x752 (config 1)
        if name in ("mobilenet_conv7", "gray_conv6"):
            sizes.append((480, 640))
        if name == "mobilenet_conv7":
            sizes.append((480, 752))
        for (h, wd) in sizes:
            imgs = synth.band_limited_images(2, h, wd, c, seed=h + c + (wd if h == 480 else 0))
            d64 = netvlad.describe(imgs, w, dtype="float64")
            gold["%s_%dx%d_desc64" % (name, h, wd)] = d64
    np.savez_compressed(os.path.join(OUT, "netvlad_golden.npz"), **gold)

    rng = np.random.default_rng(2024)
    pg = {}
    for c in range(6):
        n = [200, 200, 64, 20, 333, 150][c]
        X, uv, T, mask = dls_pnp.synth_candidate(rng, n=n, outlier_frac=[0.2, 0.0, 0.3, 0.1, 0.4, 0.2][c])
        tab = dls_pnp.sample_table(99, c, 50, n)
        r = dls_pnp.ransac_pnp(X, uv, tab)
        p2 = dls_pnp.RansacParameters(adaptive=False, max_iterations=50)
        r2 = dls_pnp.ransac_pnp(X, uv, tab, p2)
        pg["c%d_X" % c], pg["c%d_uv" % c], pg["c%d_Ttrue" % c], pg["c%d_tab" % c] = X, uv, T, tab
        for tag, rr in (("adaptive", r), ("fixed", r2)):
            pg["c%d_%s_T" % (c, tag)] = rr["T"]
            pg["c%d_%s_meta" % (c, tag)] = np.array([rr["confidence"], rr["num_iterations"], rr["n_inliers"], rr["best_hyp"], rr["best_cost"]])
    np.savez_compressed(os.path.join(OUT, "pnp_golden.npz"), **pg)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
