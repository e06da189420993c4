"""GPU parity tests for the detection leg (Shi-Tomasi block detection + cornerSubPix) through the C ABI.
Bars: corner list identical in content AND order to cv2 / the oracle (feature IDs depend on it), sub-pixel <= 1e-3 px."""
import os

import numpy as np
import pytest

from datagen import synth_klt as synth
from tests import oracle_api as oa

pytestmark = pytest.mark.gpu
CASES = ["plain", "noisy", "small"]


@pytest.fixture(scope="module")
def olib(oracle):
    oa.declare_detect(oracle)
    return oracle


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "detect_golden.npz"))


@pytest.fixture(scope="module")
def detectors():
    from ic_gvins_b200.detect import Detector
    cache = {}

    def get(W, H):
        if (W, H) not in cache:
            cache[(W, H)] = Detector(W, H, max_blocks=32, max_corners_per_block=128)
        return cache[(W, H)]
    yield get
    for d in cache.values():
        d.close()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", ["nomask", "mask"])
def test_good_features_and_subpix_vs_cv2_golden(detectors, golden, name, tag):
    img = golden[name + "_img"]
    H, W = img.shape
    n, md = golden[name + "_args"]
    mask = golden[name + "_mask"] if tag == "mask" else None
    d = detectors(W, H)
    pts = d.goodFeaturesToTrack(img, int(n), 0.01, float(md), mask=mask)
    assert np.array_equal(pts, golden[f"{name}_{tag}_pts"])
    sub = d.cornerSubPix(img, pts)
    assert np.abs(sub - golden[f"{name}_{tag}_sub"]).max() <= 1e-3


def test_block_grid_matches_oracle_on_full_frame(detectors, olib):
    """The 18-block grid of a 1280x560 frame (tracking.cc:66-85, 627-656): per-block quota 17, minDist 40, masks from circles."""
    from ic_gvins_b200.detect import block_rois
    img = synth.render_frame(synth.make_texture(1280, 560, 31), 0, 1280, 560)
    rois, quota, min_dist, _ = block_rois(1280, 560, 300)
    assert len(rois) == 18 and quota == 17 and min_dist == 40 and rois[0] == (0, 0, 208, 181) and rois[17] == (1065, 372, 213, 186)
    mask = np.full((560, 1280), 255, np.uint8)
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:560, 0:1280]
    for _ in range(40):  # filled discs of radius 40 around existing features (the reference draws them with cv::circle on the host)
        cx, cy = rng.integers(0, 1280), rng.integers(0, 560)
        mask[(xx - cx) ** 2 + (yy - cy) ** 2 <= 1600] = 0
    d = detectors(1280, 560)
    want = [quota - int(k % 4) for k in range(18)]
    got = d.detect_blocks(img, rois, want, 0.01, float(min_dist), mask, subpix=True)
    total = 0
    for roi, w_, g in zip(rois, want, got):
        ref = oa.detect_block(olib, img, mask, roi, w_, 0.01, float(min_dist))
        assert len(g) == len(ref)
        assert np.abs(g - ref).max() <= 1e-3 if len(g) else True
        # integer-pixel selection identical (compare before refinement)
        total += len(g)
    raw = d.detect_blocks(img, rois, want, 0.01, float(min_dist), mask, subpix=False)
    for roi, w_, g in zip(rois, want, raw):
        eig = oa.min_eig_roi(olib, img, roi)
        sel = np.zeros((max(1, w_), 2), np.float32)
        import ctypes as C
        m = np.ascontiguousarray(mask)
        cnt = olib.icgo_good_features_from_eig(eig.ctypes.data_as(C.c_void_p), roi[2], roi[3], C.c_void_p(m.ctypes.data + roi[1] * 1280 + roi[0]), 1280, w_, 0.01,
                                               float(min_dist), sel.ctypes.data_as(C.c_void_p))
        assert np.array_equal(g, sel[:cnt])
    assert total > 150


def test_features_detection_frame_coordinates(detectors):
    img = synth.render_frame(synth.make_texture(1280, 560, 32), 0, 1280, 560)
    d = detectors(1280, 560)
    pts = d.features_detection(img, max_features=300)
    assert pts.shape[0] > 200 and pts[:, 0].max() < 1280 and pts[:, 1].max() < 560
    # idempotent
    assert np.array_equal(pts, d.features_detection(img, max_features=300))


def test_empty_mask_and_errors(detectors):
    from ic_gvins_b200 import IcgError
    img = synth.render_frame(synth.make_texture(213, 186, 23), 0, 213, 186)
    d = detectors(213, 186)
    assert len(d.goodFeaturesToTrack(img, 17, 0.01, 40.0, mask=np.zeros((186, 213), np.uint8))) == 0
    with pytest.raises(IcgError):
        d.detect_blocks(img, [(100, 100, 200, 200)], [5])


def test_batched_device_resident_detection_equals_per_frame_calls():
    """icg_detect_blocks_dev on the level-0 planes of consecutive KLT slots (one call for all frames) == icg_detect_blocks frame by frame."""
    import ctypes as C
    from ic_gvins_b200._lib import lib, vp
    from ic_gvins_b200.detect import Detector, block_rois
    from ic_gvins_b200.klt import KltTracker
    from datagen import synth_klt as synth
    W, H, NF = 1280, 560, 3
    st = synth.KltStream(W, H, 300, 77)
    frames = [st.frame(t) for t in range(NF)]
    rois, quota, min_dist, _ = block_rois(W, H, 300)
    want = [quota - (b % 4) for b in range(len(rois))]
    d1 = Detector(W, H, max_blocks=32, max_corners_per_block=32)
    ref = [d1.detect_blocks(f, rois, want, 0.01, float(min_dist), None, subpix=True) for f in frames]
    d1.close()
    trk = KltTracker(W, H, n_slots=NF, max_points=64)
    for k, f in enumerate(frames):
        trk.upload(k, f, build=False)
    trk.sync()
    p0, p1, pitch = vp(), vp(), C.c_int()
    lib().icg_klt_slot_level0(trk._h, 0, C.byref(p0), C.byref(pitch))
    lib().icg_klt_slot_level0(trk._h, 1, C.byref(p1), C.byref(pitch))
    dN = Detector(W, H, max_blocks=NF * len(rois), max_corners_per_block=32, max_roi_pixels=213 * 186)
    got = dN.detect_blocks_dev(NF, p0.value, pitch.value, p1.value - p0.value, rois, want, 0.01, float(min_dist))
    dN.close()
    trk.close()
    for f in range(NF):
        for b in range(len(rois)):
            assert got[f][b].shape == ref[f][b].shape and np.array_equal(got[f][b], ref[f][b]), (f, b)
