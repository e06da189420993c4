"""Host-side mirror of the reference's `Camera` (IG/tracking/camera.cc): same method names and argument meaning, arithmetic in
libicgvins_b200.so (host functions: SURVEY 8a row A6 keeps the camera model on the CPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, vp


class CameraStruct(C.Structure):
    """ctypes image of `icg_camera`."""
    _fields_ = [(k, C.c_double) for k in ("fx", "fy", "cx", "cy", "skew", "k1", "k2", "p1", "p2", "k3")]


class Camera:
    def __init__(self, intrinsic, distortion):
        """Camera::createCamera (camera.cc:48-70): intrinsic = [fx, fy, cx, cy(, skew)], distortion = [k1, k2, p1, p2(, k3)]"""
        i, d = list(map(float, intrinsic)), list(map(float, distortion))
        self.c = CameraStruct(i[0], i[1], i[2], i[3], i[4] if len(i) == 5 else 0.0, d[0], d[1], d[2], d[3], d[4] if len(d) == 5 else 0.0)

    def _pts(self, pts):
        return np.ascontiguousarray(np.array(pts, np.float32).reshape(-1, 2))

    def undistortPoints(self, pts):
        p = self._pts(pts)
        check(lib().icg_camera_undistort_points(C.byref(self.c), vp(p.ctypes.data), p.shape[0]), "icg_camera_undistort_points")
        return p

    def distortPoints(self, pts):
        p = self._pts(pts)
        check(lib().icg_camera_distort_points(C.byref(self.c), vp(p.ctypes.data), p.shape[0]), "icg_camera_distort_points")
        return p

    def distortCameraPoint(self, pc):
        a = np.ascontiguousarray(np.array(pc, np.float64).reshape(-1, 3))
        out = np.zeros((a.shape[0], 2), np.float32)
        check(lib().icg_camera_distort_camera_points(C.byref(self.c), vp(a.ctypes.data), vp(out.ctypes.data), a.shape[0]), "icg_camera_distort_camera_points")
        return out

    def pixel2cam(self, pts):
        p = self._pts(pts)
        out = np.zeros((p.shape[0], 3))
        check(lib().icg_camera_pixel2cam(C.byref(self.c), vp(p.ctypes.data), vp(out.ctypes.data), p.shape[0]), "icg_camera_pixel2cam")
        return out

    def world2pixel(self, world, R, t):
        a = np.ascontiguousarray(np.array(world, np.float64).reshape(-1, 3))
        Rm, tv = np.ascontiguousarray(np.array(R, np.float64).reshape(3, 3)), np.ascontiguousarray(np.array(t, np.float64).reshape(3))
        out = np.zeros((a.shape[0], 2), np.float32)
        check(lib().icg_camera_world2pixel(C.byref(self.c), vp(Rm.ctypes.data), vp(tv.ctypes.data), vp(a.ctypes.data), vp(out.ctypes.data), a.shape[0]),
              "icg_camera_world2pixel")
        return out


def calculate_histogram(image) -> float:
    """Tracking::calculateHistigram (IG/tracking/tracking.cc:88-104): the brightness statistic of the histogram gate"""
    img = np.ascontiguousarray(image, np.uint8)
    out = C.c_double()
    check(lib().icg_tracking_histogram(vp(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], C.byref(out)), "icg_tracking_histogram")
    return out.value


def findFundamentalMat(points1, points2, ransacReprojThreshold=3.0, confidence=0.99, maxIters=1000):
    """cv2.findFundamentalMat(points1, points2, cv2.FM_RANSAC, ransacReprojThreshold, confidence) -> (F, status) as
    Tracking::trackReferenceFrame uses it (IG/tracking/tracking.cc:547)"""
    p1 = np.ascontiguousarray(np.array(points1, np.float32).reshape(-1, 2))
    p2 = np.ascontiguousarray(np.array(points2, np.float32).reshape(-1, 2))
    assert p1.shape == p2.shape
    st = np.zeros(p1.shape[0], np.uint8)
    F = np.zeros(9)
    check(lib().icg_find_fundamental_mat_ransac(vp(p1.ctypes.data), vp(p2.ctypes.data), p1.shape[0], float(ransacReprojThreshold), float(confidence),
                                                int(maxIters), vp(st.ctypes.data), vp(F.ctypes.data)), "icg_find_fundamental_mat_ransac")
    return F.reshape(3, 3), st


def triangulatePoints(Tcw0, Tcw1, pc0, pc1):
    """Tracking::triangulatePoint (IG/tracking/tracking.cc:796-808) for n pairs: Tcw0 (n, 3, 4), Tcw1 (3, 4), pc0 / pc1 (n, 2 or 3) -> pw (n, 3)"""
    a = np.ascontiguousarray(np.array(Tcw0, np.float64).reshape(-1, 12))
    b = np.ascontiguousarray(np.array(Tcw1, np.float64).reshape(12))
    p0 = np.ascontiguousarray(np.array(pc0, np.float64).reshape(a.shape[0], -1)[:, :2])
    p1 = np.ascontiguousarray(np.array(pc1, np.float64).reshape(a.shape[0], -1)[:, :2])
    out = np.zeros((a.shape[0], 3))
    check(lib().icg_triangulate_points(vp(a.ctypes.data), vp(b.ctypes.data), vp(p0.ctypes.data), vp(p1.ctypes.data), a.shape[0], vp(out.ctypes.data)),
          "icg_triangulate_points")
    return out
