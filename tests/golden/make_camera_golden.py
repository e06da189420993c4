"""Generates tests/golden/camera_golden.npz with cv2 (the OpenCV the reference links): `cv::undistortPoints(pts, pts, K, D, Mat(), K)`
exactly as Camera::undistortPoints calls it (ic_gvins/ic_gvins/tracking/camera.cc:72-74).  Run in the build container (cv2 4.13.0)."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CAMS = {
    # name: (fx, fy, cx, cy, skew), (k1, k2, p1, p2, k3), (W, H)
    "mild_1280x560": ((787.0, 786.1, 640.2, 281.5, 0.0), (-0.0512, 0.0421, 0.0007, -0.0004, 0.0), (1280, 560)),
    "strong_skew": ((787.0, 786.1, 640.2, 281.5, 0.3), (-0.31, 0.12, 0.0007, -0.0004, -0.02), (1280, 560)),
    "fisheye_like_640x480": ((380.0, 379.5, 320.5, 241.0, 0.0), (-0.28, 0.07, 0.0002, 0.0001, 0.0), (640, 480)),
}


def main():
    rng = np.random.default_rng(7)
    out = {}
    for name, (intr, dist, (W, H)) in CAMS.items():
        K = np.array([[intr[0], intr[4], intr[2]], [0, intr[1], intr[3]], [0, 0, 1]])
        D = np.array(dist)
        pts = np.stack([rng.uniform(-5, W + 5, 300), rng.uniform(-5, H + 5, 300)], 1).astype(np.float32)
        und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
        out[name + "_intr"], out[name + "_dist"], out[name + "_pts"], out[name + "_undist"] = np.array(intr), D, pts, und
    np.savez_compressed(os.path.join(HERE, "camera_golden.npz"), cv2_version=cv2.__version__, **out)
    print("wrote camera_golden.npz:", list(CAMS))


if __name__ == "__main__":
    main()
