"""Generates tests/golden/fundamental_golden.npz with cv2 (the OpenCV the reference links):
`cv::findFundamentalMat(pts1, pts2, cv::FM_RANSAC, reprojection_error_std_, 0.99, status)` exactly as Tracking::trackReferenceFrame
calls it (ic_gvins/ic_gvins/tracking/tracking.cc:547).  Run in the build container (cv2 4.13.0)."""
import math
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(rng, n, n_out, noise, yaw, trans, f=787.0, c=(640.0, 280.0)):
    P = np.stack([rng.uniform(-10, 10, n), rng.uniform(-4, 4, n), rng.uniform(8, 40, n)], 1)
    R = np.array([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
    p1 = P[:, :2] / P[:, 2:] * f + np.array(c)
    Q = (R @ P.T).T + np.array(trans)
    p2 = Q[:, :2] / Q[:, 2:] * f + np.array(c)
    p1 += rng.normal(0, noise, p1.shape)
    p2 += rng.normal(0, noise, p2.shape)
    out = rng.choice(n, n_out, replace=False)
    p2[out] += rng.uniform(-30, 30, (n_out, 2))
    return p1.astype(np.float32), p2.astype(np.float32)


def main():
    rng = np.random.default_rng(99)
    cases = {}
    spec = [("n300_std1p5", 300, 60, 0.3, 0.03, (0.5, 0.05, 0.1), 1.5), ("n120_std1p5", 120, 25, 0.3, 0.10, (0.4, 0.0, 0.2), 1.5),
            ("n40_clean", 40, 0, 0.05, 0.05, (0.3, 0.1, 0.0), 1.0), ("n15_minimal", 15, 3, 0.2, 0.08, (0.6, -0.1, 0.1), 1.5),
            ("n200_heavy_outliers", 200, 110, 0.4, 0.06, (0.5, 0.0, 0.3), 2.0), ("n100_tight", 100, 20, 0.5, 0.02, (0.2, 0.02, 0.05), 0.5)]
    for name, n, n_out, noise, yaw, trans, thr in spec:
        p1, p2 = scene(rng, n, n_out, noise, yaw, trans)
        F, st = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, thr, 0.99)
        cases[name + "_p1"], cases[name + "_p2"], cases[name + "_thr"] = p1, p2, np.array([thr])
        cases[name + "_status"] = st.ravel().astype(np.uint8)
        cases[name + "_F"] = F
    np.savez_compressed(os.path.join(HERE, "fundamental_golden.npz"), cv2_version=cv2.__version__, **cases)
    print("wrote fundamental_golden.npz:", [s[0] for s in spec])


if __name__ == "__main__":
    main()
