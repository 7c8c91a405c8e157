"""Micro-benchmark of the correspondence front end: a batch of loop candidates, 5000 ORB-like features per image
(cv::ORB::create(5000), PointFeatureMatching.cpp:17), device time of matcher + GMS from the library's CUDA events, host
call time, and the CPU reference beside it (cv2.BFMatcher + the compiled reference GMS matcher when present)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cerebro_b200.frontend import FrontEnd  # noqa: E402
from oracle import gms  # noqa: E402
from tests.test_frontend import _synthetic_pair  # noqa: E402


def main():
    n_pairs, nf, w, h = 16, 5000, 640, 480
    rng = np.random.default_rng(0)
    pairs = [_synthetic_pair(rng, nf, nf, w, h) for _ in range(n_pairs)]
    fe = FrontEnd(max_pairs=n_pairs, max_features=nf)
    args = ([p[0] for p in pairs], [p[1] for p in pairs], [p[2] for p in pairs], [p[3] for p in pairs], (w, h), (w, h))
    for _ in range(3):
        fe.match_gms(*args)
    dev, host = [], []
    for _ in range(10):
        t0 = time.perf_counter()
        res = fe.match_gms(*args)
        host.append(time.perf_counter() - t0)
        dev.append(fe.last_match_ms())
    dev_ms, host_ms = float(np.median(dev)), float(np.median(host)) * 1e3
    pair_evals = n_pairs * nf * nf
    out = {"pairs": n_pairs, "features": nf, "device_ms_match_gms": round(dev_ms, 3), "host_call_ms": round(host_ms, 3),
           "pairs_per_s_device": round(n_pairs / dev_ms * 1e3, 1), "descriptor_comparisons_per_s": pair_evals / (dev_ms * 1e-3),
           "popc_per_s": 8 * pair_evals / (dev_ms * 1e-3), "gms_inliers_pair0": res[0]["n_inliers"]}
    try:
        import cv2

        t0 = time.perf_counter()
        m = cv2.BFMatcher(cv2.NORM_HAMMING).match(pairs[0][1], pairs[0][3])
        t1 = time.perf_counter()
        idx = np.array([x.trainIdx for x in m], dtype=np.int32)
        if gms.reference_available():
            gms.gms_reference(pairs[0][0], (w, h), pairs[0][2], (w, h), np.arange(nf), idx)
        t2 = time.perf_counter()
        out["cpu_reference_ms_per_pair"] = {"cv2_BFMatcher": round((t1 - t0) * 1e3, 2), "reference_gms": round((t2 - t1) * 1e3, 2),
                                            "cores": os.cpu_count(), "threads_opencv": cv2.getNumThreads()}
    except Exception as e:  # noqa: BLE001
        out["cpu_reference_ms_per_pair"] = "unavailable: %s" % e
    print(json.dumps(out))
    fe.close()


if __name__ == "__main__":
    main()
