"""Micro-benchmark of the image side of a loop candidate at 480 x 640: cv::remap, cv::ORB(5000, FAST threshold 0) and
cv::StereoBM(64, 21) + the 3-D image on the device (library CUDA-event times + blocking host call times), with the installed
OpenCV timed beside them on the box's host cores when cv2 is importable.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200.features import Features  # noqa: E402
from cerebro_b200.frontend import FrontEnd  # noqa: E402
from tests.synth_orb import image, rect_maps  # noqa: E402
from tests.synth_stereo import stereo_scene  # noqa: E402


def med(f, n=7, warm=2):
    for _ in range(warm):
        f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    h, w, nimg, npair = 480, 640, 2, 8
    imgs = np.stack([image("textured", h, w, 20 + i) for i in range(nimg)])
    mx, my = rect_maps(h, w, 5)
    out = {"rows": h, "cols": w}
    ft = Features(h, w, max_images=nimg, max_keypoints=6000)
    ft.set_remap(0, mx, my)
    out["remap_host_call_ms_per_%d_images" % nimg] = round(med(lambda: ft.remap(imgs, 0)), 3)
    dev = []
    host = med(lambda: (ft.orb(imgs, 5000), dev.append(ft.last_orb_ms)))
    out["orb5000_host_call_ms_per_%d_images" % nimg] = round(host, 3)
    out["orb5000_stream_ms_per_%d_images" % nimg] = round(float(np.median(dev[2:])), 3)
    out["orb_keypoints"] = [int(len(k["pt"])) for k in ft.orb(imgs, 5000)]
    ft.close()
    pairs = [stereo_scene(h, w, i % 3, 30 + i) for i in range(npair)]
    L = np.stack([p[0] for p in pairs])
    R = np.stack([p[1] for p in pairs])
    fe = FrontEnd(max_pairs=npair, max_features=5000)
    dev = []
    host = med(lambda: (fe.stereo_bm(L, R), dev.append(fe.last_stereo_ms())))
    out["stereo_bm_host_call_ms_per_%d_pairs" % npair] = round(host, 3)
    out["stereo_bm_device_ms_per_pair"] = round(float(np.median(dev[2:])) / npair, 4)
    d = fe.stereo_bm(L, R)
    Q = np.array([[1, 0, 0, -w / 2], [0, 1, 0, -h / 2], [0, 0, 0, 400.0], [0, 0, 1 / 0.11, 0]])
    out["disparity_to_3d_host_call_ms_per_%d_pairs" % npair] = round(med(lambda: fe.disparity_to_3d(d, Q)), 3)
    fe.close()
    try:
        import cv2

        orb = cv2.ORB_create(5000)
        orb.setFastThreshold(0)
        bm = cv2.StereoBM_create(64, 21)
        out["cpu_opencv_ms"] = {
            "remap_per_image": round(med(lambda: cv2.remap(imgs[0], mx, my, cv2.INTER_LINEAR)), 3),
            "orb5000_per_image": round(med(lambda: orb.detectAndCompute(imgs[0], None), n=5, warm=1), 2),
            "stereo_bm_per_pair": round(med(lambda: bm.compute(L[0], R[0]), n=5, warm=1), 2),
            "threads_opencv": cv2.getNumThreads(), "cores": os.cpu_count()}
    except Exception as e:  # noqa: BLE001
        out["cpu_opencv_ms"] = "unavailable: %s" % e
    print(json.dumps(out))


if __name__ == "__main__":
    main()
