"""SURVEY 8a row A6 -- the camera model.  The product functions are HOST code inside libicgvins_b200.so (no GPU needed), so these parity
tests run in the CPU suite: cv::undistortPoints against the cv2 golden vectors (float outputs bit-identical), the in-repo radtan /
projection arithmetic against the numpy restatement in oracle/camera_ref.py (which follows camera.cc line by line)."""
import os

import numpy as np
import pytest

from oracle import camera_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "camera_golden.npz")


def cases():
    g = np.load(GOLD)
    return sorted(k[:-5] for k in g.files if k.endswith("_intr"))


def cam_dict(intr, dist):
    return dict(fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], skew=intr[4], k1=dist[0], k2=dist[1], p1=dist[2], p2=dist[3], k3=dist[4])


@pytest.mark.parametrize("name", cases())
def test_undistort_points_matches_cv2_golden(name):
    from ic_gvins_b200.camera import Camera
    g = np.load(GOLD)
    intr, dist, pts, und = g[name + "_intr"], g[name + "_dist"], g[name + "_pts"], g[name + "_undist"]
    assert np.array_equal(ref.undistort_points(cam_dict(intr, dist), pts), und)        # the restatement is pinned ...
    assert np.array_equal(Camera(intr, dist).undistortPoints(pts), und)                # ... and the product matches bit for bit


@pytest.mark.parametrize("name", cases())
def test_radtan_and_projection_match_the_restatement(name):
    from ic_gvins_b200.camera import Camera
    g = np.load(GOLD)
    intr, dist, pts = g[name + "_intr"], g[name + "_dist"], g[name + "_pts"]
    cd, cam = cam_dict(intr, dist), Camera(intr, dist)
    assert np.array_equal(cam.distortPoints(pts), ref.distort_points(cd, pts))
    assert np.array_equal(cam.pixel2cam(pts), ref.pixel2cam(cd, pts))
    rng = np.random.default_rng(3)
    pc = np.stack([rng.uniform(-20, 20, 200), rng.uniform(-8, 8, 200), rng.uniform(4, 60, 200)], 1)
    assert np.array_equal(cam.distortCameraPoint(pc), ref.distort_camera_point(cd, pc))
    a = 0.3
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0.0]])
    t = np.array([1.0, -2.0, 0.5])
    pw = (R @ pc.T).T + t
    assert np.abs(cam.world2pixel(pw, R, t) - ref.world2pixel(cd, pw, R, t)).max() <= 1e-4  # summation order of the 3x3 product may differ by 1 ulp


def test_distort_inverts_undistort():
    """round trip inside the image: undistort then distort returns the pixel to ~1e-3 px (the five fixed iterations are not exact)"""
    from ic_gvins_b200.camera import Camera
    g = np.load(GOLD)
    intr, dist, pts = g["mild_1280x560_intr"], g["mild_1280x560_dist"], g["mild_1280x560_pts"]
    cam = Camera(intr, dist)
    back = cam.distortPoints(cam.undistortPoints(pts))
    assert np.abs(back - pts).max() < 2e-3


def test_camera_rejects_bad_arguments():
    import ctypes as C
    from ic_gvins_b200._lib import lib
    from ic_gvins_b200.camera import CameraStruct
    c = CameraStruct(0.0, 1.0, 0, 0, 0, 0, 0, 0, 0, 0)  # fx = 0
    p = np.zeros((4, 2), np.float32)
    assert lib().icg_camera_undistort_points(C.byref(c), C.c_void_p(p.ctypes.data), 4) != 0
    assert b"bad arguments" in lib().icg_last_error()


def test_histogram_gate_statistic_matches_cv2():
    """Tracking::calculateHistigram (tracking.cc:88-104) restated with cv2.calcHist + the reference's float / double sequence"""
    cv2 = pytest.importorskip("cv2")
    from ic_gvins_b200.camera import calculate_histogram
    rng = np.random.default_rng(11)
    for img in (rng.integers(0, 256, (560, 1280), dtype=np.uint8), np.full((480, 640), 200, np.uint8),
                (rng.normal(120, 30, (217, 333)).clip(0, 255)).astype(np.uint8)):
        h = cv2.calcHist([img], [0], None, [256], [0, 256]).reshape(-1)  # float32 counts
        acc = 0.0
        for k in range(256):
            acc += float(np.float32(h[k]) * np.float32(k)) / 256.0
        assert calculate_histogram(img) == acc / (img.shape[0] * img.shape[1])
