"""GPU parity tests for CLAHE (icg_clahe_apply / _apply_dev) through the C ABI: bit-exact against the cv2 golden vectors and the oracle."""
import os

import numpy as np
import pytest

from tests import oracle_api as oa

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "clahe_golden.npz")


def cases():
    g = np.load(GOLD)
    return sorted(k[:-3] for k in g.files if k.endswith("_in"))


@pytest.mark.parametrize("name", cases())
def test_clahe_matches_cv2_golden(name):
    from ic_gvins_b200.clahe import Clahe
    g = np.load(GOLD)
    img, ref = g[name + "_in"], g[name + "_out"]
    tx, ty, clip = g[name + "_par"]
    c = Clahe(img.shape[1], img.shape[0], clip, (int(tx), int(ty)))
    out = c.apply(img)
    assert np.array_equal(out, ref)
    inplace = img.copy()
    c.apply(inplace, inplace)  # clahe_->apply(image, image) as the reference calls it (tracking.cc:141)
    assert np.array_equal(inplace, ref)
    c.close()


def test_clahe_matches_oracle_on_stream_frames(oracle):
    """the bench's synthetic frames, at the reference's size and parameters (1280 x 560, clip 3.0, 21 x 21 tiles)"""
    from datagen import synth_klt
    from ic_gvins_b200.clahe import Clahe
    st = synth_klt.KltStream(1280, 560, 10, 77)
    c = Clahe(1280, 560, 3.0, (21, 21))
    for t in range(3):
        img = st.frame(t)
        assert np.array_equal(c.apply(img), oa.clahe_apply(oracle, img, 3.0, 21, 21))
    c.close()


def test_clahe_device_resident_into_klt_slot(oracle):
    """upload a raw frame into a KLT slot, equalise it in place on the device, build the pyramid: level 0 must equal the oracle's CLAHE"""
    import ctypes as C
    from ic_gvins_b200._lib import check, lib, vp
    from ic_gvins_b200.clahe import Clahe
    from ic_gvins_b200.klt import KltTracker
    rng = np.random.default_rng(5)
    W, H = 640, 480
    img = rng.integers(0, 256, (H, W), dtype=np.uint8)
    img[100:200, 50:300] //= 4
    trk = KltTracker(W, H, n_slots=2, max_points=16)
    trk.upload(0, img, build=False)
    ptr, pitch = vp(), C.c_int()
    check(lib().icg_klt_slot_level0(trk._h, 0, C.byref(ptr), C.byref(pitch)), "icg_klt_slot_level0")
    trk.sync()
    c = Clahe(W, H, 3.0, (21, 21))
    c.apply_dev(ptr.value, pitch.value, ptr.value, pitch.value)
    c.sync()
    trk.build_pyramids(0, 1)
    got = trk.download_level(0, 0)
    assert np.array_equal(got, oa.clahe_apply(oracle, img, 3.0, 21, 21))
    c.close()
    trk.close()
