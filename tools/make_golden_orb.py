"""Generates tests/golden/orb_golden.npz in the build container from the INSTALLED OpenCV (the reference's own dependency):
seeded images -> cv2.ORB_create(n) + setFastThreshold(0) + detectAndCompute (src/utils/PointFeatureMatching.cpp:16-22), and
cv2.remap with CV_32FC1 maps (src/utils/CameraGeometry.cpp:42, 381-382).  The GPU tests compare the device against these
fixtures (and against cv2 itself when it is importable)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.synth_orb import CASES, REMAP_CASES, image, rect_maps  # noqa: E402


def main():
    out = {}
    for name, h, w, n, kind, seed in CASES:
        img = image(kind, h, w, seed)
        orb = cv2.ORB_create(n)
        orb.setFastThreshold(0)
        k, d = orb.detectAndCompute(img, None)
        out[name + "_kps"] = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in k], np.float32)
        out[name + "_desc"] = d
        out[name + "_n"] = np.array([n])
    for name, h, w, seed in REMAP_CASES:
        img = image("textured", h, w, seed)
        m1, m2 = rect_maps(h, w, seed), rect_maps(h, w, seed + 100)
        und = cv2.remap(img, m1[0], m1[1], cv2.INTER_LINEAR)
        rec = cv2.remap(und, m2[0], m2[1], cv2.INTER_LINEAR)
        out[name + "_undistorted"] = und
        out[name + "_rectified"] = rec
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "orb_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
