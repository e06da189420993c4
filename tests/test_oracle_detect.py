"""Pin the detection oracle (oracle/detect_ref.c) against cv2-generated golden vectors.  CPU only.
Bars: eig map bit-exact (CRC), corner list identical in content and order, sub-pixel positions <= 1e-3 px."""
import zlib

import numpy as np
import pytest

from datagen import synth_klt as synth
from tests import oracle_api as oa

CASES = ["plain", "noisy", "small"]


@pytest.fixture(scope="module")
def olib(oracle):
    oa.declare_detect(oracle)
    return oracle


@pytest.fixture(scope="module")
def golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "detect_golden.npz"))


@pytest.mark.parametrize("name", CASES)
def test_min_eig_bit_exact(olib, golden, name):
    img = golden[name + "_img"]
    H, W = img.shape
    eig = oa.min_eig_roi(olib, img, (0, 0, W, H))
    assert zlib.crc32(eig.tobytes()) == int(golden[name + "_eig_crc"][0])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", ["nomask", "mask"])
def test_good_features_identical_list_and_order(olib, golden, name, tag):
    img = golden[name + "_img"]
    n, md = golden[name + "_args"]
    mask = golden[name + "_mask"] if tag == "mask" else None
    pts = oa.good_features(olib, img, int(n), 0.01, float(md), mask)
    assert np.array_equal(pts, golden[f"{name}_{tag}_pts"])
    sub = oa.corner_subpix(olib, img, pts)
    assert np.abs(sub - golden[f"{name}_{tag}_sub"]).max() <= 1e-3


def test_roi_semantics_match_cv2_primitives(olib, golden):
    """C++ ROI view: Sobel reads the parent frame beyond the block edge, the covariance box filter reflects at the block edge."""
    img = synth.render_frame(synth.make_texture(1280, 560, 31), 0, 1280, 560)
    assert zlib.crc32(img.tobytes()) == int(golden["roi_crc"][0])
    x0, y0, w, h = [int(v) for v in golden["roi_rect"]]
    eig = oa.min_eig_roi(olib, img, (x0, y0, w, h))
    # The golden map was composed from cv2.Sobel on the 1280-wide frame (all columns on the AVX2/FMA row path).  In C++ the row
    # filter runs over the 208-wide ROI, so its last 208 % 32 = 16 columns take the scalar non-FMA path (see detect_ref.c):
    # the composition is exact for the columns whose 3x3 support lies in the FMA part.
    fma_cols = (w & ~31) - 1
    assert np.array_equal(eig[:, :fma_cols], golden["roi_eig"][:, :fma_cols])
    # and it differs from the isolated-copy semantics in the border ring only
    iso = oa.min_eig_roi(olib, np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w]), (0, 0, w, h))
    assert np.array_equal(iso[2:-2, 2:-2], eig[2:-2, 2:-2]) and not np.array_equal(iso, eig)
