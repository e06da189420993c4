"""oracle/camera_ref.py -- numpy FP64 restatement of the reference's camera model (TEST INFRASTRUCTURE ONLY).

Follows ic_gvins/ic_gvins/tracking/camera.cc line by line: pixel2cam (:126-130), cam2pixel (:132-134), distortPoints (:76-90),
distortCameraPoint (:106-120), world2cam / world2pixel (:144-150).  undistortPoints (:72-74) forwards to cv::undistortPoints, which is
un-vendored: its published algorithm (five fixed-point iterations; skew ignored when normalising, P = K applied with its skew) is
restated in `undistort_points` and pinned against cv2 4.13.0 by tests/golden/camera_golden.npz."""
import numpy as np


def pixel2cam(cam, px):
    px = np.asarray(px, np.float32).reshape(-1, 2).astype(np.float64)
    y = (px[:, 1] - cam["cy"]) / cam["fy"]
    x = (px[:, 0] - cam["cx"] - cam["skew"] * y) / cam["fx"]
    return np.stack([x, y, np.ones_like(x)], 1)


def cam2pixel(cam, pc):
    pc = np.asarray(pc, np.float64).reshape(-1, 3)
    u = (cam["fx"] * pc[:, 0] + cam["skew"] * pc[:, 1]) / pc[:, 2] + cam["cx"]
    v = cam["fy"] * pc[:, 1] / pc[:, 2] + cam["cy"]
    return np.stack([u, v], 1).astype(np.float32)


def _radtan(cam, x, y):
    r2 = x * x + y * y
    rr = 1 + cam["k1"] * r2 + cam["k2"] * r2 * r2 + cam["k3"] * r2 * r2 * r2
    return x * rr + 2 * cam["p1"] * x * y + cam["p2"] * (r2 + 2 * x * x), y * rr + cam["p1"] * (r2 + 2 * y * y) + 2 * cam["p2"] * x * y


def distort_points(cam, px):
    pc = pixel2cam(cam, px)
    xd, yd = _radtan(cam, pc[:, 0], pc[:, 1])
    return cam2pixel(cam, np.stack([xd, yd, np.ones_like(xd)], 1))


def distort_camera_point(cam, pc):
    pc = np.asarray(pc, np.float64).reshape(-1, 3)
    xd, yd = _radtan(cam, pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2])
    xd, yd = xd.astype(np.float32).astype(np.float64), yd.astype(np.float32).astype(np.float64)  # static_cast<float> (:114-115)
    return cam2pixel(cam, np.stack([xd, yd, np.ones_like(xd)], 1))


def world2pixel(cam, pw, R, t):
    pw = np.asarray(pw, np.float64).reshape(-1, 3)
    return cam2pixel(cam, (np.asarray(R, np.float64).T @ (pw - np.asarray(t, np.float64)).T).T)


def undistort_points(cam, px, iters=5):
    px = np.asarray(px, np.float32).reshape(-1, 2).astype(np.float64)
    x = (px[:, 0] - cam["cx"]) * (1.0 / cam["fx"])
    y = (px[:, 1] - cam["cy"]) * (1.0 / cam["fy"])
    x0, y0 = x.copy(), y.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1 + ((cam["k3"] * r2 + cam["k2"]) * r2 + cam["k1"]) * r2)
        dx = 2 * cam["p1"] * x * y + cam["p2"] * (r2 + 2 * x * x)
        dy = cam["p1"] * (r2 + 2 * y * y) + 2 * cam["p2"] * x * y
        x, y = (x0 - dx) * icdist, (y0 - dy) * icdist
    return np.stack([cam["fx"] * x + cam["skew"] * y + cam["cx"], cam["fy"] * y + cam["cy"]], 1).astype(np.float32)
