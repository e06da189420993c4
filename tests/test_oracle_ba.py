"""Self-consistency known-answer tests for the BA oracle (oracle/ba_ref.cpp).  The reference ships no tests and Ceres is
absent (parity unpinned), so the oracle is pinned by: finite-difference Jacobians through
PoseParameterization::Plus, the IMU factor vanishing on noise-free self-consistent data, and LM convergence.  CPU only."""
import copy

import numpy as np
import pytest

from datagen import synth_ba
from tests import oracle_api as oa


@pytest.fixture(scope="module")
def olib(oracle):
    oa.declare_ba(oracle)
    return oracle


@pytest.fixture(scope="module")
def window(olib):
    return synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=10, L=300, seed=2024)


def fd_pose(fun, x, eps, olib):
    cols = []
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        rp = fun(oa.pose_plus(olib, x, d))
        rm = fun(oa.pose_plus(olib, x, -d))
        cols.append((rp - rm) / (2 * eps))
    return np.stack(cols, axis=1)


def test_reprojection_jacobians_fd(olib, window):
    prob, _ = window
    pose = prob["pose"].reshape(-1, 7)
    for f in (0, 5, 77, 400):
        c = prob["f_const"][14 * f:14 * f + 14]
        i, j, l = prob["f_ref"][f], prob["f_obs"][f], prob["f_lm"][f]
        args = [pose[i].copy(), pose[j].copy(), prob["ext"][:7].copy(), prob["invdepth"][l], prob["ext"][7]]
        r, Js = oa.reproj_eval(olib, *args, c, prob["reproj_std"])
        for w in range(3):
            def fun(x, w=w):
                a = list(args); a[w] = x
                return oa.reproj_eval(olib, *a, c, prob["reproj_std"], False)[0]
            Jfd = fd_pose(fun, args[w], 1e-6, olib)
            assert np.abs(Jfd - Js[w][:, :6]).max() <= 1e-5 * max(1.0, np.abs(Jfd).max())
            assert np.all(Js[w][:, 6] == 0)
        for w, eps in ((3, 1e-7), (4, 1e-6)):
            a = list(args); a[w] = args[w] + eps
            b = list(args); b[w] = args[w] - eps
            Jfd = (oa.reproj_eval(olib, *a, c, prob["reproj_std"], False)[0] - oa.reproj_eval(olib, *b, c, prob["reproj_std"], False)[0]) / (2 * eps)
            assert np.abs(Jfd - Js[w].ravel()).max() <= 1e-5 * max(1.0, np.abs(Jfd).max())


def test_imu_jacobians_fd(olib, window):
    prob, _ = window
    pose, mix = prob["pose"].reshape(-1, 7), prob["mix"].reshape(-1, 9)
    off, pn = prob["pn_off"], prob["pn"].reshape(-1, 4)
    k = 3
    blob = prob["imu_blob"][480 * k:480 * (k + 1)]
    pnk = pn[off[k]:off[k + 1]]
    args = [pose[k].copy(), mix[k].copy(), pose[k + 1].copy(), mix[k + 1].copy()]
    r, Js = oa.imu_eval(olib, blob, pnk, *args)
    for w in (0, 2):
        def fun(x, w=w):
            a = list(args); a[w] = x
            return oa.imu_eval(olib, blob, pnk, *a, False)[0]
        Jfd = fd_pose(fun, args[w], 1e-6, olib)
        # the reference's attitude blocks are first-order (small-residual) approximations: compare at 1e-3 relative
        assert np.abs(Jfd - Js[w][:, :6]).max() <= 2e-3 * np.abs(Jfd).max()
    for w in (1, 3):
        cols = []
        for c in range(9):
            eps = 1e-6 if c < 3 else 1e-9
            a = list(args); a[w] = args[w].copy(); a[w][c] += eps
            b = list(args); b[w] = args[w].copy(); b[w][c] -= eps
            cols.append((oa.imu_eval(olib, blob, pnk, *a, False)[0] - oa.imu_eval(olib, blob, pnk, *b, False)[0]) / (2 * eps))
        Jfd = np.stack(cols, axis=1)
        assert np.abs(Jfd - Js[w]).max() <= 2e-3 * np.abs(Jfd).max()


def test_imu_residual_small_at_truth(olib):
    prob, truth = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=6, L=40, seed=3, perturb=False)
    pose, mix = prob["pose"].reshape(-1, 7), prob["mix"].reshape(-1, 9)
    off, pn = prob["pn_off"], prob["pn"].reshape(-1, 4)
    for k in range(5):
        r, _ = oa.imu_eval(olib, prob["imu_blob"][480 * k:480 * (k + 1)], pn[off[k]:off[k + 1]], pose[k], mix[k], pose[k + 1], mix[k + 1], False)
        assert np.abs(r).max() < 6.0  # whitened residual of a consistent (noisy) trajectory is O(1)


def test_lm_reduces_cost_and_recovers_truth(olib, window):
    prob, truth = window
    p = copy.deepcopy(prob)
    p["ext_const"], p["td_const"] = 1, 1
    p["ext"] = truth["ext"].copy()
    s = oa.ba_solve(olib, p, 20)
    assert s["final_cost"] < 1e-3 * s["initial_cost"]
    assert np.abs(p["pose"].reshape(-1, 7)[:, :3] - truth["pose"][:, :3]).max() < 0.25  # GNSS sigma 0.05-0.1 m
    assert np.median(np.abs(p["invdepth"] / truth["invdepth"] - 1)) < 0.02


def test_residual_costs_and_culling_protocol(olib, window):
    prob, _ = window
    p = copy.deepcopy(prob)
    # inject gross outliers in a few observations
    fc = p["f_const"].reshape(-1, 14)
    fc[10, 3] += 0.2
    fc[500, 4] -= 0.15
    oa.ba_solve(olib, p, 5)
    rc, gc = oa.ba_residual_costs(olib, p)
    assert 2 * rc[10] > 5.991 and 2 * rc[500] > 5.991
    assert (2 * rc > 5.991).sum() < 0.05 * len(rc)


@pytest.fixture(scope="module")
def window_normal(olib):
    return synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=6, L=60, seed=77, earth=False)


def test_preintegration_normal_jacobians_fd(olib, window_normal):
    """PreintegrationNormal (iswithearth false; preintegration_normal.cc:38-153): blob tagged, residual / Jacobians vs finite differences."""
    prob, _ = window_normal
    pose, mix = prob["pose"].reshape(-1, 7), prob["mix"].reshape(-1, 9)
    k = 2
    blob = prob["imu_blob"][480 * k:480 * (k + 1)]
    assert blob[477] == 1.0 and np.all(blob[20:27] == 0)
    none = np.zeros((0, 4))
    args = [pose[k].copy(), mix[k].copy(), pose[k + 1].copy(), mix[k + 1].copy()]
    r, Js = oa.imu_eval(olib, blob, none, *args)
    for w in (0, 2):
        def fun(x, w=w):
            a = list(args); a[w] = x
            return oa.imu_eval(olib, blob, none, *a, False)[0]
        Jfd = fd_pose(fun, args[w], 1e-6, olib)
        assert np.abs(Jfd - Js[w][:, :6]).max() <= 2e-3 * np.abs(Jfd).max()
        assert np.all(Js[w][:, 6] == 0)
    for w in (1, 3):
        cols = []
        for c in range(9):
            eps = 1e-6 if c < 3 else 1e-9
            a = list(args); a[w] = args[w].copy(); a[w][c] += eps
            b = list(args); b[w] = args[w].copy(); b[w][c] -= eps
            cols.append((oa.imu_eval(olib, blob, none, *a, False)[0] - oa.imu_eval(olib, blob, none, *b, False)[0]) / (2 * eps))
        Jfd = np.stack(cols, axis=1)
        assert np.abs(Jfd - Js[w]).max() <= 2e-3 * np.abs(Jfd).max()


def test_preintegration_normal_consistent_and_solvable(olib):
    prob, truth = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=6, L=60, seed=78, earth=False, perturb=False)
    pose, mix = prob["pose"].reshape(-1, 7), prob["mix"].reshape(-1, 9)
    for k in range(5):
        r, _ = oa.imu_eval(olib, prob["imu_blob"][480 * k:480 * (k + 1)], np.zeros((0, 4)), pose[k], mix[k], pose[k + 1], mix[k + 1], False)
        assert np.abs(r).max() < 6.0
    prob2, truth2 = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=6, L=60, seed=78, earth=False)
    prob2["ext_const"], prob2["td_const"] = 1, 1
    prob2["ext"] = truth2["ext"].copy()
    s = oa.ba_solve(olib, prob2, 20)
    assert s["final_cost"] < 1e-3 * s["initial_cost"]
    assert np.abs(prob2["pose"].reshape(-1, 7)[:, :3] - truth2["pose"][:, :3]).max() < 0.25
