"""SURVEY.md 8f ranks 2-4 on the device (csrc/geom.cu) vs their oracles: the cv2 golden vectors (undistortPoints, findFundamentalMat inlier
masks), the numpy restatements (radtan, triangulation) and the oracle's preintegration."""
import os

import numpy as np
import pytest

from oracle import camera_ref as cref
from oracle import fundamental_ref as fref
from tests import oracle_api as oa

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def geom():
    from ic_gvins_b200.geom import Geometry
    g = Geometry()
    yield g
    g.close()


def test_undistort_distort_on_device_match_cv2_golden(geom):
    g = np.load(os.path.join(GOLD, "camera_golden.npz"))
    for name in sorted(k[:-5] for k in g.files if k.endswith("_intr")):
        intr, dist, pts, und = g[name + "_intr"], g[name + "_dist"], g[name + "_pts"], g[name + "_undist"]
        assert np.array_equal(geom.undistortPoints(intr, dist, pts), und), name          # float outputs bit-identical to cv2
        cd = dict(fx=intr[0], fy=intr[1], cx=intr[2], cy=intr[3], skew=intr[4], k1=dist[0], k2=dist[1], p1=dist[2], p2=dist[3], k3=dist[4])
        assert np.array_equal(geom.distortPoints(intr, dist, pts), cref.distort_points(cd, pts)), name
    # a big batch (the throughput-mode use): every point of 296 frames x 300 points
    rng = np.random.default_rng(1)
    intr, dist = g["mild_1280x560_intr"], g["mild_1280x560_dist"]
    big = np.stack([rng.uniform(0, 1280, 88800), rng.uniform(0, 560, 88800)], 1).astype(np.float32)
    from ic_gvins_b200.camera import Camera
    assert np.array_equal(geom.undistortPoints(intr, dist, big), Camera(intr, dist).undistortPoints(big))


def test_ransac_on_device_inlier_masks_identical_to_cv2_golden(geom):
    g = np.load(os.path.join(GOLD, "fundamental_golden.npz"))
    for name in sorted(k[:-3] for k in g.files if k.endswith("_p1")):
        p1, p2, thr, st, F = g[name + "_p1"], g[name + "_p2"], float(g[name + "_thr"][0]), g[name + "_status"], g[name + "_F"]
        Fg, sg = geom.findFundamentalMat(p1, p2, thr, 0.99)
        assert np.array_equal(sg, st), name
        assert np.abs(Fg - F).max() <= 1e-9 * np.abs(F).max(), name


def test_ransac_on_device_equals_host_function_on_random_scenes(geom):
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLD, "make_fundamental_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mk)
    except ImportError:
        pytest.skip("cv2 not importable (the scene generator lives in the golden script)")
    from ic_gvins_b200.camera import findFundamentalMat
    rng = np.random.default_rng(2026)
    for trial in range(12):
        n = int(rng.integers(15, 300))
        p1, p2 = mk.scene(rng, n, int(0.25 * n), 0.3, float(rng.uniform(0.0, 0.15)), (0.5, float(rng.uniform(-0.1, 0.1)), 0.1))
        Fh, sh = findFundamentalMat(p1, p2, 1.5, 0.99)
        Fg, sg = geom.findFundamentalMat(p1, p2, 1.5, 0.99)
        assert np.array_equal(sg, sh), trial
        assert np.array_equal(sg, fref.find_fm_ransac(p1, p2, 1.5, 0.99).astype(np.uint8)), trial
        assert np.abs(Fg - Fh).max() <= 1e-9 * max(1e-300, np.abs(Fh).max())
    q = np.ones((20, 2), np.float32)  # every subset collinear: no model
    F, st = geom.findFundamentalMat(q, q, 1.5, 0.99)
    assert st.sum() == 0 and not F.any()


def test_triangulation_on_device(geom):
    rng = np.random.default_rng(8)
    n = 600

    def tcw(yaw, t):
        R = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        return np.hstack([R.T, (-R.T @ np.asarray(t))[:, None]])

    T1 = tcw(0.12, (0.9, 0.05, 0.2))
    T0 = np.stack([tcw(0.01 * (k % 5), (0.1 * (k % 3), 0.0, 0.0)) for k in range(n)])
    pw = np.stack([rng.uniform(-10, 10, n), rng.uniform(-4, 4, n), rng.uniform(8, 60, n)], 1)
    pc0 = np.stack([(T0[k] @ np.append(pw[k], 1.0)) for k in range(n)])
    pc1 = (T1 @ np.hstack([pw, np.ones((n, 1))]).T).T
    pc0, pc1 = pc0[:, :2] / pc0[:, 2:], pc1[:, :2] / pc1[:, 2:]
    got = geom.triangulatePoints(T0, T1, pc0, pc1)
    assert np.abs(got - pw).max() <= 1e-8 * np.abs(pw).max()
    noisy0, noisy1 = pc0 + rng.normal(0, 1e-3, pc0.shape), pc1 + rng.normal(0, 1e-3, pc1.shape)
    got = geom.triangulatePoints(T0, T1, noisy0, noisy1)
    want = np.stack([fref.triangulate_point(T0[k], T1, noisy0[k], noisy1[k]) for k in range(n)])
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()


@pytest.mark.parametrize("earth", [True, False])
def test_batched_preintegration_on_device_matches_oracle(geom, oracle, earth):
    from datagen import synth_ba
    oa.declare_ba(oracle)
    rng = np.random.default_rng(5)
    states, imus, want_blob, want_end = [], [], [], []
    for k in range(19):
        t0 = 0.5 * k
        imu = synth_ba.imu_samples(t0, t0 + 0.5 - 0.005 * (k % 3), 200.0, rng, np.zeros(3), np.zeros(3), earth=earth)
        p, v, _, psi = synth_ba.trajectory(t0)
        st = np.concatenate([p, synth_ba.q_yaw(psi), v, [1e-4, -2e-4, 3e-4], [1e-3, 2e-3, -1e-3]])
        b, _, e = oa.preintegrate(oracle, st, synth_ba.IEWN if earth else None, synth_ba.GRAVITY, synth_ba.NOISE5, imu)
        states.append(st), imus.append(imu), want_blob.append(b), want_end.append(e)
    blobs, ends = geom.imu_preintegrate_batch(np.array(states), synth_ba.IEWN if earth else None, synth_ba.GRAVITY, synth_ba.NOISE5, imus)
    for k in range(19):
        bo, bg = want_blob[k], blobs[k]
        assert bg[477] == bo[477]
        assert np.abs(bg[:27] - bo[:27]).max() <= 1e-12 * max(1.0, np.abs(bo[:27]).max())
        assert np.abs(bg[27:252] - bo[27:252]).max() <= 1e-12 * np.abs(bo[27:252]).max()
        assert np.abs(bg[252:477] - bo[252:477]).max() <= 1e-10 * np.abs(bo[252:477]).max()
        assert np.abs(ends[k] - want_end[k]).max() <= 1e-13 * np.abs(want_end[k]).max()
