"""Host-side mirror of the reference's KLT call sites (IG/tracking/tracking.cc:385-403, 487-506).

`KltTracker.calcOpticalFlowPyrLK` keeps OpenCV's argument list and return convention (as cv2 exposes it), so the
parity tests read like calls into the reference's own dependency.  Everything runs in libicgvins_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, vp

OPTFLOW_USE_INITIAL_FLOW = 4
TERM_COUNT, TERM_EPS = 1, 2


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


class KltTracker:
    """One tracker handle == one image geometry on one GPU (the reference has one Tracking object per camera)."""

    def __init__(self, width: int, height: int, n_slots: int = 4, max_points: int = 4096, device: int = 0, stream=None):
        self.W, self.H, self.n_slots, self.max_points = width, height, n_slots, max_points
        self._h = vp()
        check(lib().icg_klt_create(C.byref(self._h), width, height, n_slots, max_points, device,
                                   vp(stream) if stream else None), "icg_klt_create")

    def close(self):
        if getattr(self, "_h", None):
            lib().icg_klt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ drop-in: cv2.calcOpticalFlowPyrLK
    def calcOpticalFlowPyrLK(self, prevImg, nextImg, prevPts, nextPts, winSize=(21, 21), maxLevel=3,
                             criteria=(TERM_COUNT + TERM_EPS, 30, 0.01), flags=0):
        prevImg = np.ascontiguousarray(prevImg, dtype=np.uint8)
        nextImg = np.ascontiguousarray(nextImg, dtype=np.uint8)
        if prevImg.shape != (self.H, self.W) or nextImg.shape != (self.H, self.W):
            raise ValueError("image size does not match the tracker")
        p = np.ascontiguousarray(prevPts, dtype=np.float32).reshape(-1, 2)
        n = p.shape[0]
        if flags & OPTFLOW_USE_INITIAL_FLOW:
            q = np.array(nextPts, dtype=np.float32).reshape(-1, 2).copy()
        else:
            q = p.copy()
        status = np.zeros(n, np.uint8)
        err = np.zeros(n, np.float32)
        max_iter = criteria[1] if criteria[0] & TERM_COUNT else 30
        eps = criteria[2] if criteria[0] & TERM_EPS else 0.0
        check(lib().icg_klt_calc_optical_flow_pyr_lk(self._h, _ptr(prevImg), _ptr(nextImg), self.W, _ptr(p), _ptr(q),
                                                     _ptr(status), _ptr(err), n, winSize[0], maxLevel, max_iter,
                                                     float(eps), flags), "icg_klt_calc_optical_flow_pyr_lk")
        return q, status, err

    # ------------------------------------------------------------------ fused fwd+bwd+gates (tracking.cc:385-403)
    def track_fb(self, prevImg, nextImg, prevPts, predicted):
        prevImg = np.ascontiguousarray(prevImg, dtype=np.uint8)
        nextImg = np.ascontiguousarray(nextImg, dtype=np.uint8)
        p = np.ascontiguousarray(prevPts, dtype=np.float32).reshape(-1, 2)
        q = np.array(predicted, dtype=np.float32).reshape(-1, 2).copy()
        n = p.shape[0]
        back = np.zeros((n, 2), np.float32)
        status = np.zeros(n, np.uint8)
        check(lib().icg_klt_track_fb(self._h, _ptr(prevImg), _ptr(nextImg), self.W, _ptr(p), _ptr(q), _ptr(back),
                                     _ptr(status), n), "icg_klt_track_fb")
        return q, back, status

    # ------------------------------------------------------------------ device-resident API
    def upload(self, slot: int, img: np.ndarray, build: bool = True):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        f = lib().icg_klt_upload if build else lib().icg_klt_upload_level0
        check(f(self._h, slot, _ptr(img), img.strides[0]), "icg_klt_upload")

    def upload_ptr(self, slot: int, host_ptr: int, stride: int, build: bool = True):
        f = lib().icg_klt_upload if build else lib().icg_klt_upload_level0
        check(f(self._h, slot, vp(host_ptr), stride), "icg_klt_upload")

    def upload_batch_ptrs(self, first_slot: int, host_ptrs, stride: int):
        """One call for the new frame of every stream: host_ptrs = iterable of host addresses (pinned), slots first_slot ...; no pyramid build."""
        arr = (C.c_void_p * len(host_ptrs))(*host_ptrs)
        check(lib().icg_klt_upload_batch(self._h, first_slot, len(host_ptrs), arr, stride), "icg_klt_upload_batch")

    def build_pyramids(self, first_slot: int, count: int):
        check(lib().icg_klt_build_pyramids(self._h, first_slot, count), "icg_klt_build_pyramids")

    def level_shape(self, level: int):
        w, h, pitch, ptr = C.c_int(), C.c_int(), C.c_int(), vp()
        check(lib().icg_klt_slot_level(self._h, 0, level, C.byref(ptr), C.byref(pitch), C.byref(w), C.byref(h)), "slot_level")
        return h.value, w.value

    def download_level(self, slot: int, level: int) -> np.ndarray:
        h, w = self.level_shape(level)
        out = np.zeros((h, w), np.uint8)
        check(lib().icg_klt_download_level(self._h, slot, level, _ptr(out), w), "icg_klt_download_level")
        return out

    def track_batch_dev(self, n_total: int, slots_ptr: int, prev_ptr: int, init_ptr: int, fwd_ptr: int, bwd_ptr: int,
                        status_ptr: int, mode: int = 1):
        check(lib().icg_klt_track_batch_dev(self._h, n_total, vp(slots_ptr), vp(prev_ptr), vp(init_ptr), vp(fwd_ptr),
                                            vp(bwd_ptr) if bwd_ptr else None, vp(status_ptr), mode), "icg_klt_track_batch_dev")

    def sync(self):
        check(lib().icg_klt_sync(self._h), "icg_klt_sync")
