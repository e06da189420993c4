"""SURVEY 8f rank 3 (first half): the fundamental-matrix RANSAC gate of Tracking::trackReferenceFrame (tracking.cc:546-555).  The product
function is HOST code inside libicgvins_b200.so, so the parity tests run in the CPU suite: inlier masks identical to the cv2 golden
vectors, for the numpy restatement (oracle/fundamental_ref.py) and for the product; the model itself to 1e-9."""
import os

import numpy as np
import pytest

from oracle import fundamental_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fundamental_golden.npz")


def cases():
    g = np.load(GOLD)
    return sorted(k[:-3] for k in g.files if k.endswith("_p1"))


@pytest.mark.parametrize("name", cases())
def test_inlier_mask_identical_to_cv2_golden(name):
    from ic_gvins_b200.camera import findFundamentalMat
    g = np.load(GOLD)
    p1, p2, thr, st, F = g[name + "_p1"], g[name + "_p2"], float(g[name + "_thr"][0]), g[name + "_status"], g[name + "_F"]
    assert np.array_equal(ref.find_fm_ransac(p1, p2, thr, 0.99), st.astype(bool))  # the restatement is pinned ...
    Fg, sg = findFundamentalMat(p1, p2, thr, 0.99)                                 # ... and the product matches
    assert np.array_equal(sg, st)
    assert np.abs(Fg - F).max() <= 1e-9 * np.abs(F).max()


def test_random_scenes_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(__file__), "golden", "make_fundamental_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    from ic_gvins_b200.camera import findFundamentalMat
    rng = np.random.default_rng(2025)
    for trial in range(10):
        n = int(rng.integers(15, 300))
        p1, p2 = mk.scene(rng, n, int(0.25 * n), 0.3, float(rng.uniform(0.0, 0.15)), (0.5, float(rng.uniform(-0.1, 0.1)), 0.1))
        _, st = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.5 / 787.0 * 787.0, 0.99)
        _, sg = findFundamentalMat(p1, p2, 1.5, 0.99)
        assert np.array_equal(sg, st.ravel()), trial


def test_degenerate_inputs():
    from ic_gvins_b200._lib import IcgError
    from ic_gvins_b200.camera import findFundamentalMat
    p = np.zeros((10, 2), np.float32)
    with pytest.raises(IcgError, match="at least 15"):
        findFundamentalMat(p, p)
    # all points identical: every subset is collinear -> no model, empty mask (cv2 returns None / zeros)
    q = np.ones((20, 2), np.float32)
    F, st = findFundamentalMat(q, q, 1.5, 0.99)
    assert st.sum() == 0 and not F.any()


def test_triangulation_matches_the_restatement_and_recovers_the_points():
    """Tracking::triangulatePoint: product vs the numpy SVD restatement (1e-9 relative), and ground truth on noise-free views"""
    from ic_gvins_b200.camera import triangulatePoints
    rng = np.random.default_rng(8)
    n = 200

    def tcw(yaw, t):
        R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])  # R_w_c
        return np.hstack([R.T, (-R.T @ np.asarray(t))[:, None]])

    T1 = tcw(0.12, (0.9, 0.05, 0.2))
    T0 = np.stack([tcw(0.01 * (k % 5), (0.1 * (k % 3), 0.0, 0.0)) for k in range(n)])
    pw = np.stack([rng.uniform(-10, 10, n), rng.uniform(-4, 4, n), rng.uniform(8, 60, n)], 1)
    pc0 = np.stack([(T0[k] @ np.append(pw[k], 1.0)) for k in range(n)])
    pc1 = (T1 @ np.hstack([pw, np.ones((n, 1))]).T).T
    pc0, pc1 = pc0[:, :2] / pc0[:, 2:], pc1[:, :2] / pc1[:, 2:]
    got = triangulatePoints(T0, T1, pc0, pc1)
    assert np.abs(got - pw).max() <= 1e-8 * np.abs(pw).max()
    noisy0, noisy1 = pc0 + rng.normal(0, 1e-3, pc0.shape), pc1 + rng.normal(0, 1e-3, pc1.shape)
    got = triangulatePoints(T0, T1, noisy0, noisy1)
    want = np.stack([ref.triangulate_point(T0[k], T1, noisy0[k], noisy1[k]) for k in range(n)])
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
