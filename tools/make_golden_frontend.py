"""Generates tests/golden/gms_golden.npz in the build container: two synthetic image pairs -> cv2.ORB (FAST threshold 0)
-> cv2.BFMatcher(NORM_HAMMING).match -> the UNMODIFIED reference GMS matcher (oracle/_ref/libgms_ref.so, built by
oracle/Makefile from /root/reference/src/utils/GMSMatcher).  These outputs come from the reference's own code / its own
OpenCV dependency, not from the oracle restatement."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gms  # noqa: E402


def textured(rng, h, w, sigma):
    img = (rng.random((h, w)) * 255).astype(np.uint8)
    img = cv2.GaussianBlur(img, (0, 0), sigma)
    return cv2.normalize(img, None, 0, 255, cv2.NORM_MINMAX)


def main():
    assert gms.reference_available(), "run `make -C oracle` first"
    rng = np.random.default_rng(2024)
    out = {}
    cases = [
        ("easy", 480, 640, 2.0, np.array([[1.02, 0.03, 8], [-0.02, 0.99, 5], [1e-5, -2e-5, 1]]), 3000),
        ("hard", 480, 752, 1.5, np.array([[0.93, -0.12, 60], [0.10, 0.95, -25], [6e-5, 3e-5, 1]]), 2000),
    ]
    for name, h, w, sigma, H, nfeat in cases:
        a = textured(rng, h, w, sigma)
        b = cv2.warpPerspective(a, H, (w, h))
        b = cv2.add(b, (rng.random((h, w)) * 12).astype(np.uint8))
        orb = cv2.ORB_create(nfeat)
        orb.setFastThreshold(0)
        k1, d1 = orb.detectAndCompute(a, None)
        k2, d2 = orb.detectAndCompute(b, None)
        m = cv2.BFMatcher(cv2.NORM_HAMMING).match(d1, d2)
        q = np.array([x.queryIdx for x in m], dtype=np.int32)
        t = np.array([x.trainIdx for x in m], dtype=np.int32)
        dist = np.array([x.distance for x in m], dtype=np.int32)
        assert np.array_equal(q, np.arange(len(k1)))
        kp1 = np.array([k.pt for k in k1], dtype=np.float32)
        kp2 = np.array([k.pt for k in k2], dtype=np.float32)
        mask, n = gms.gms_reference(kp1, (w, h), kp2, (w, h), q, t)
        print(name, len(k1), len(k2), "gms inliers", n)
        for key, val in (("kp1", kp1), ("kp2", kp2), ("d1", d1), ("d2", d2), ("train", t), ("dist", dist), ("mask", mask),
                         ("size", np.array([w, h]))):
            out["%s_%s" % (name, key)] = val
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gms_golden.npz"), **out)


if __name__ == "__main__":
    main()
