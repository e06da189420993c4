"""Host-side mirror of `cv::CLAHE` as the reference uses it (IG/tracking/tracking.cc:62,141): `Clahe(W, H, clipLimit, tileGridSize)`
plays `cv2.createCLAHE(clipLimit, tileGridSize)`, `.apply(img)` is `clahe->apply(img, img)`.  All arithmetic runs in libicgvins_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib, vp


class Clahe:
    def __init__(self, width: int, height: int, clipLimit: float = 3.0, tileGridSize=(21, 21), device: int = 0, stream=None):
        self.W, self.H = width, height
        self._h = vp()
        check(lib().icg_clahe_create(C.byref(self._h), width, height, int(tileGridSize[0]), int(tileGridSize[1]), float(clipLimit), device,
                                     vp(stream) if stream else None), "icg_clahe_create")

    def close(self):
        if getattr(self, "_h", None):
            lib().icg_clahe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, src, dst=None):
        """cv2.CLAHE.apply(src[, dst]) on an 8-bit single-channel image; pass dst=src for the reference's in-place call."""
        src = np.ascontiguousarray(src, np.uint8) if dst is not src else src
        assert src.shape == (self.H, self.W)
        out = np.empty_like(src) if dst is None else dst
        check(lib().icg_clahe_apply(self._h, vp(src.ctypes.data), src.strides[0], vp(out.ctypes.data), out.strides[0]), "icg_clahe_apply")
        return out

    def apply_dev(self, dev_src: int, src_pitch: int, dev_dst: int, dst_pitch: int):
        """device-resident variant (asynchronous): raw device pointers, e.g. a KLT slot's level-0 plane"""
        check(lib().icg_clahe_apply_dev(self._h, vp(dev_src), src_pitch, vp(dev_dst), dst_pitch), "icg_clahe_apply_dev")

    def apply_batch_dev(self, n_frames: int, dev_src: int, src_pitch: int, src_frame_stride: int, dev_dst: int, dst_pitch: int, dst_frame_stride: int,
                        want_hist: bool = False):
        """n_frames device-resident frames in one launch pair (in place allowed).  want_hist: also returns Tracking::calculateHistigram of every RAW
        frame (the histogram-gate statistic, accumulated by the LUT pass) as an (n_frames,) float64 array; the call then synchronises."""
        hist = np.zeros(n_frames) if want_hist else None
        check(lib().icg_clahe_apply_batch_dev(self._h, n_frames, vp(dev_src), src_pitch, src_frame_stride, vp(dev_dst), dst_pitch, dst_frame_stride,
                                              vp(hist.ctypes.data) if want_hist else None), "icg_clahe_apply_batch_dev")
        return hist

    def sync(self):
        check(lib().icg_clahe_sync(self._h), "icg_clahe_sync")
