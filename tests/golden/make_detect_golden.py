"""Golden vectors for the detection leg from cv2 4.13.0 (run HERE; needs cv2):  python tests/golden/make_detect_golden.py

Stand-alone calls exactly as Tracking::featuresDetection makes them on a block (tracking.cc:647,651) but on whole small
frames (cv2 from Python cannot express a C++ ROI view with a live parent, see SURVEY.md 8c "ROI hazard"):
  cv2.goodFeaturesToTrack(img, n, 0.01, minDist, mask) ; cv2.cornerSubPix(img, pts, (5,5), (-1,-1), (COUNT+EPS, 20, 0.01))
plus cv2.cornerMinEigenVal CRCs and, for the ROI semantics, the eig map of a block composed from cv2 primitives
(cv2.Sobel on the full frame -> products -> cv2.boxFilter on the ROI-sized covariance, normalize=False).
"""
import os
import sys
import zlib

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from datagen import synth_klt as synth  # noqa: E402

CRIT = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.01)


def frame(seed, W, H, noise=0):
    img = synth.render_frame(synth.make_texture(W, H, seed), 0, W, H)
    if noise:
        rng = np.random.default_rng(seed)
        img = np.clip(img.astype(int) + rng.integers(-noise, noise, img.shape), 0, 255).astype(np.uint8)
    return img


def roi_eig_from_primitives(img, x0, y0, w, h):
    s = 1.0 / (4 * 3 * 255.0)
    dx = cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=3, scale=s)[y0:y0 + h, x0:x0 + w]
    dy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=3, scale=s)[y0:y0 + h, x0:x0 + w]
    cov = np.ascontiguousarray(np.stack([dx * dx, dx * dy, dy * dy], axis=2))
    box = cv2.boxFilter(cov, -1, (3, 3), normalize=False, borderType=cv2.BORDER_DEFAULT)
    a, b, c = box[..., 0] * np.float32(0.5), box[..., 1], box[..., 2] * np.float32(0.5)
    return ((a + c) - np.sqrt((a - c) * (a - c) + b * b)).astype(np.float32)


def main():
    out = {}
    cases = {"plain": (21, 640, 360, 0, 60, 40.0), "noisy": (22, 480, 270, 25, 80, 20.0), "small": (23, 213, 186, 0, 17, 40.0)}
    for name, (seed, W, H, noise, n, md) in cases.items():
        img = frame(seed, W, H, noise)
        mask = np.full((H, W), 255, np.uint8)
        cv2.circle(mask, (W // 2, H // 2), 40, 0, cv2.FILLED)
        cv2.circle(mask, (W // 5, H // 4), 40, 0, cv2.FILLED)
        out[name + "_img"] = img
        out[name + "_mask"] = mask
        out[name + "_args"] = np.array([n, md], np.float64)
        eig = cv2.cornerMinEigenVal(img, 3, ksize=3)
        out[name + "_eig_crc"] = np.array([zlib.crc32(eig.tobytes())], np.uint64)
        for tag, m in (("nomask", None), ("mask", mask)):
            pts = cv2.goodFeaturesToTrack(img, n, 0.01, md, mask=m).reshape(-1, 2)
            sub = cv2.cornerSubPix(img, pts.reshape(-1, 1, 2).copy(), (5, 5), (-1, -1), CRIT).reshape(-1, 2)
            out[f"{name}_{tag}_pts"] = pts
            out[f"{name}_{tag}_sub"] = sub
            print(name, tag, len(pts), "corners")
    # ROI semantics: block 7 of the 1280x560 grid (213x186 blocks, shrunk by 5) on a 1280x560 frame
    img = frame(31, 1280, 560, 0)
    x0, y0, w, h = 213, 186, 208, 181
    out["roi_crc"] = np.array([zlib.crc32(img.tobytes())], np.uint64)
    out["roi_rect"] = np.array([x0, y0, w, h], np.int32)
    out["roi_eig"] = roi_eig_from_primitives(img, x0, y0, w, h)
    path = os.path.join(os.path.dirname(__file__), "detect_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
