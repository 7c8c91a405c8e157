"""Generates tests/golden/stereo_golden.npz in the build container: small synthetic rectified pairs -> the installed
OpenCV's cv2.StereoBM (the reference's own dependency, src/utils/CameraGeometry.cpp:81) -> int16 disparities."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.synth_stereo import stereo_scene  # noqa: E402


def main():
    out = {}
    for name, (h, w, nd, ws, kind, seed) in {"a": (120, 200, 32, 9, 0, 1), "b": (97, 211, 64, 21, 1, 2), "c": (150, 260, 64, 21, 2, 3)}.items():
        left, right = stereo_scene(h, w, kind, seed)
        out[name + "_disp"] = cv2.StereoBM_create(nd, ws).compute(left, right)
        out[name + "_cfg"] = np.array([h, w, nd, ws, kind, seed])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stereo_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
