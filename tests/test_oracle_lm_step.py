"""Known-answer test of the oracle's trust-region step against an INDEPENDENT dense numpy restatement of Ceres' published algorithm
("Schur-vs-full-normal-equation solve equality", "Huber corrector vs closed form").  The oracle eliminates the inverse
depths with a Schur complement and solves the reduced camera system; here the full Jacobian of the same window is assembled from the
oracle's single-factor evaluations, robustified with the closed-form Huber corrector, Jacobi-scaled, damped with the clamped
Levenberg-Marquardt diagonal and solved DENSE -- the accepted point after one iteration must agree.  CPU only."""
import copy
import math

import numpy as np
import pytest

from datagen import synth_ba
from tests import oracle_api as oa


@pytest.fixture(scope="module")
def olib(oracle):
    oa.declare_ba(oracle)
    return oracle


GB_STD = 7200 / 3600.0 * math.pi / 180.0  # imu_error_factor.h:89-91
AB_STD = 2.0e4 * 1.0e-5


def dense_system(olib, prob):
    """residual vector r, dense Jacobian J (local coordinates), cost; columns [pose_k 6 | mix_k 9]_k, ext 6, td 1, rho_l"""
    K, L = prob["K"], prob["L"]
    pose, mix, ext = prob["pose"].reshape(-1, 7), prob["mix"].reshape(-1, 9), prob["ext"]
    col_pose = lambda k: 15 * k
    col_mix = lambda k: 15 * k + 6
    col_ext, col_td, col_rho = 15 * K, 15 * K + 6, 15 * K + 7
    n = 15 * K + 7 + L
    rows_r, rows_J, cost = [], [], 0.0

    def add(r, blocks, huber):
        nonlocal cost
        r = np.asarray(r, float)
        J = np.zeros((len(r), n))
        for c0, Jb in blocks:
            J[:, c0:c0 + Jb.shape[1]] += Jb
        s = float(r @ r)
        if huber and s > 1.0:
            # HuberLoss(1): rho = 2 sqrt(s) - 1, rho' = 1 / sqrt(s), rho'' < 0 -> Corrector: r, J scaled by sqrt(rho') (alpha = 0)
            cost += 0.5 * (2.0 * math.sqrt(s) - 1.0)
            sc = math.sqrt(1.0 / math.sqrt(s))
            r, J = r * sc, J * sc
        else:
            cost += 0.5 * s
        rows_r.append(r)
        rows_J.append(J)

    off, pn = prob["pn_off"], prob["pn"].reshape(-1, 4)
    for g, nd in enumerate(prob["gnss_node"]):
        r = np.zeros(3)
        J = np.zeros((3, 7))
        a = [pose[nd].copy(), prob["gnss_blh"][3 * g:3 * g + 3].copy(), prob["gnss_std"][3 * g:3 * g + 3].copy(), np.array(prob["lever"], np.float64)]
        olib.icgo_gnss_eval(oa._p(a[0]), oa._p(a[1]), oa._p(a[2]), oa._p(a[3]), oa._p(r), oa._p(J))
        add(r, [(col_pose(int(nd)), J[:, :6])], bool(prob["gnss_huber"]))
    for k in range(prob["n_imu"]):
        r, Js = oa.imu_eval(olib, prob["imu_blob"].reshape(-1, 480)[k], pn[off[k]:off[k + 1]], pose[k], mix[k], pose[k + 1], mix[k + 1])
        add(r, [(col_pose(k), Js[0][:, :6]), (col_mix(k), Js[1]), (col_pose(k + 1), Js[2][:, :6]), (col_mix(k + 1), Js[3])], False)
    if prob["has_imu_error"]:
        k = prob["n_imu"]
        r = np.concatenate([mix[k][3:6] / GB_STD, mix[k][6:9] / AB_STD])
        J = np.zeros((6, 9))
        for q in range(3):
            J[q, 3 + q], J[3 + q, 6 + q] = 1.0 / GB_STD, 1.0 / AB_STD
        add(r, [(col_mix(k), J)], False)
    for f in range(prob["F"]):
        if not prob["f_active"][f]:
            continue
        i, j, l = int(prob["f_ref"][f]), int(prob["f_obs"][f]), int(prob["f_lm"][f])
        r, Js = oa.reproj_eval(olib, pose[i], pose[j], ext[:7], prob["invdepth"][l], ext[7], prob["f_const"][14 * f:14 * f + 14], prob["reproj_std"])
        add(r, [(col_pose(i), Js[0][:, :6]), (col_pose(j), Js[1][:, :6]), (col_ext, Js[2][:, :6]), (col_rho + l, Js[3]), (col_td, Js[4])],
            bool(prob["reproj_huber"]))
    return np.concatenate(rows_r), np.vstack(rows_J), cost


def apply_step(olib, prob, delta):
    K, L = prob["K"], prob["L"]
    q = copy.deepcopy(prob)
    pose, mix = q["pose"].reshape(-1, 7), q["mix"].reshape(-1, 9)
    for k in range(K):
        pose[k] = oa.pose_plus(olib, pose[k], delta[15 * k:15 * k + 6])
        mix[k] += delta[15 * k + 6:15 * k + 15]
    q["ext"][:7] = oa.pose_plus(olib, q["ext"][:7].copy(), delta[15 * K:15 * K + 6])
    q["ext"][7] += delta[15 * K + 6]
    q["invdepth"] = q["invdepth"] + delta[15 * K + 7:]
    q["pose"], q["mix"] = pose.reshape(-1), mix.reshape(-1)
    return q


@pytest.mark.parametrize("seed,huber", [(31, True), (32, False)])
def test_one_lm_iteration_equals_the_dense_normal_equations(olib, seed, huber):
    prob = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=5, L=30, seed=seed, pixel_noise=2.0 if huber else 0.5)[0]
    prob["reproj_huber"], prob["gnss_huber"] = int(huber), int(huber)
    r, J, cost0 = dense_system(olib, prob)
    if huber:
        pose = prob["pose"].reshape(-1, 7)
        assert any(float(np.sum(oa.reproj_eval(olib, pose[prob["f_ref"][f]], pose[prob["f_obs"][f]], prob["ext"][:7], prob["invdepth"][prob["f_lm"][f]],
                                               prob["ext"][7], prob["f_const"][14 * f:14 * f + 14], prob["reproj_std"], False)[0] ** 2)) > 1.0
                   for f in range(prob["F"])), "the case must exercise the robust branch"
    # Ceres LM: Jacobi scaling 1 / (1 + ||col||), D^2 = clamp(diag(J'^T J'), 1e-6, 1e32) / radius, radius_0 = 1e4
    radius = 1e4
    s = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    Js = J * s[None, :]
    H = Js.T @ Js
    D2 = np.clip(np.diag(H), 1e-6, 1e32) / radius
    g = Js.T @ r
    step_s = np.linalg.solve(H + np.diag(D2), -g)
    delta = step_s * s
    model_change = -(step_s @ g) - 0.5 * step_s @ H @ step_s  # -(J' step)^T (r + J' step / 2)
    cand = apply_step(olib, prob, delta)
    cost1 = dense_system(olib, cand)[2]
    rho = (cost0 - cost1) / model_change
    assert rho > 1e-3, "the test window must produce an accepted first step"
    # the oracle: same window, ONE iteration of Solver::Solve (Schur-eliminated inverse depths, reduced camera system)
    po = copy.deepcopy(prob)
    so = oa.ba_solve(olib, po, 1)
    assert so["iterations"] == 1 and so["num_successful_steps"] >= 1
    assert abs(so["initial_cost"] - cost0) <= 1e-12 * cost0
    assert abs(so["final_cost"] - cost1) <= 1e-9 * cost1
    for key in ("pose", "mix", "ext", "invdepth"):
        assert np.abs(po[key] - cand[key]).max() <= 1e-9 * max(1.0, np.abs(cand[key]).max()), key
