"""Pin of the oracle's WHOLE trust-region trajectory (not one step) against an independent dense numpy restatement of Ceres'
TrustRegionMinimizer + LevenbergMarquardtStrategy loop (SURVEY.md 8c "Ceres semantics"), on the cfg-3 window with the reference's
two-pass protocol (IG/ic_gvins.cc:1130-1239): 5 iterations with Huber on GNSS + reprojection, chi-square culling, 15 iterations with
GNSS un-robustified.

Independent = no Schur complement, no reduced camera system, no C++ loop: the full (3 300 x 457) Jacobian is assembled from
single-factor evaluations, robustified with the closed-form Huber corrector, Jacobi-scaled once, damped with the clamped LM diagonal
and solved DENSE with numpy; step acceptance, radius update and the three tolerances follow Ceres' published loop.  Compared per
iteration: accepted / rejected, trust-region radius, cost; at the end: iteration count, termination and the solution.  The oracle's
per-iteration state is read by re-solving from the same start with max_num_iterations = 1, 2, ... (its summary carries the iteration
count, the number of successful steps, the cost and the radius).  CPU only."""
import copy
import math

import numpy as np
import pytest

from datagen import synth_ba
from tests import oracle_api as oa
from tests.test_oracle_lm_step import AB_STD, GB_STD, apply_step  # noqa: F401  (same column layout)


@pytest.fixture(scope="module")
def olib(oracle):
    oa.declare_ba(oracle)
    return oracle


def quat_mul(a, b):  # xyzw
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz])


def dense_system(olib, prob, want_jac=True):
    """(r, J, cost) of the whole window with every factor type the solve handles (GNSS, IMU, bias-magnitude, first-window priors,
    marginalization prior, reprojection); columns [pose_k 6 | mix_k 9]_k, ext 6, td 1, rho_l.  Constant blocks get zero columns."""
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
    if prob["has_pose_prior"]:
        r = np.zeros(6)
        J = np.zeros((6, 7))
        a = [pose[0].copy(), np.asarray(prob["pose_prior"], float).copy(), np.asarray(prob["pose_prior_std"], float).copy()]
        olib.icgo_pose_prior_eval(oa._p(a[0]), oa._p(a[1]), oa._p(a[2]), oa._p(r), oa._p(J))
        add(r, [(col_pose(0), J[:, :6])], False)
    if prob["has_mix_prior"]:  # ImuMixPriorFactor (imu_mix_prior_factor.h:40-75): r = (mix - prior) / std, J = diag(1 / std)
        sd = np.asarray(prob["mix_prior_std"], float)
        add((mix[0] - np.asarray(prob["mix_prior"], float)) / sd, [(col_mix(0), np.diag(1.0 / sd))], False)
    if prob["marg_r"] > 0:  # MarginalizationFactor (marginalization_factor.h:47-101): e = e0 + J0 dx, dx in local coordinates
        rr = prob["marg_r"]
        J0 = np.asarray(prob["marg_J0"], float).reshape(rr, rr)
        dx, cols, xo = [], [], 0
        x0 = np.asarray(prob["marg_x0"], float)
        for t, nd in zip(prob["marg_block_type"], prob["marg_block_node"]):
            if t in (0, 2):
                x = pose[nd] if t == 0 else ext[:7]
                xl = x0[xo:xo + 7]
                dq = quat_mul(np.array([-xl[3], -xl[4], -xl[5], xl[6]]) / (xl[3:7] @ xl[3:7]), x[3:7])
                a = 2.0 * dq[:3] * (1.0 if dq[3] >= 0 else -1.0)
                dx += list(x[:3] - xl[:3]) + list(a)
                cols += list(range(col_pose(int(nd)), col_pose(int(nd)) + 6)) if t == 0 else list(range(col_ext, col_ext + 6))
                xo += 7
            elif t == 1:
                dx += list(mix[nd] - x0[xo:xo + 9])
                cols += list(range(col_mix(int(nd)), col_mix(int(nd)) + 9))
                xo += 9
            else:
                dx.append(ext[7] - x0[xo])
                cols.append(col_td)
                xo += 1
        r = np.asarray(prob["marg_e0"], float) + J0 @ np.array(dx)
        J = np.zeros((rr, n))
        J[:, cols] = J0
        cost += 0.5 * float(r @ r)
        rows_r.append(r)
        rows_J.append(J)
    for f in range(prob["F"]):
        if not prob["f_active"][f]:
            continue
        i, j, l = int(prob["f_ref"][f]), int(prob["f_obs"][f]), int(prob["f_lm"][f])
        r, Js = oa.reproj_eval(olib, pose[i], pose[j], ext[:7], prob["invdepth"][l], ext[7], prob["f_const"][14 * f:14 * f + 14], prob["reproj_std"],
                               want_jac)
        add(r, [(col_pose(i), Js[0][:, :6]), (col_pose(j), Js[1][:, :6]), (col_ext, Js[2][:, :6]), (col_rho + l, Js[3]), (col_td, Js[4])],
            bool(prob["reproj_huber"]))
    r, J = np.concatenate(rows_r), np.vstack(rows_J)
    if prob["ext_const"]:
        J[:, col_ext:col_ext + 6] = 0
    if prob["td_const"]:
        J[:, col_td] = 0
    return r, J, cost


def active_mask(prob):
    """Parameter columns Ceres keeps in the reduced program: constant blocks and inverse depths no active factor touches are removed."""
    K, L = prob["K"], prob["L"]
    m = np.ones(15 * K + 7 + L, bool)
    if prob["ext_const"]:
        m[15 * K:15 * K + 6] = False
    if prob["td_const"]:
        m[15 * K + 6] = False
    used = np.zeros(L, bool)
    used[np.asarray(prob["f_lm"])[np.asarray(prob["f_active"]) != 0]] = True
    m[15 * K + 7:] = used
    return m


def x_norm(prob, mask):
    K = prob["K"]
    s = float(prob["pose"] @ prob["pose"] + prob["mix"] @ prob["mix"])
    if not prob["ext_const"]:
        s += float(prob["ext"][:7] @ prob["ext"][:7])
    if not prob["td_const"]:
        s += float(prob["ext"][7] ** 2)
    rho = prob["invdepth"][mask[15 * K + 7:]]
    return math.sqrt(s + float(rho @ rho))


def dense_lm(olib, prob, max_iter):
    """Ceres TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy, dense.  Returns (solved problem, trace): trace[k] =
    (successful, radius after the iteration, cost after the iteration) for every executed iteration, + termination."""
    P = copy.deepcopy(prob)
    mask = active_mask(P)
    r, J, cost = dense_system(olib, P)
    scale = np.where(mask, 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0))), 0.0)  # jacobi_scaling, computed once
    radius, decrease, invalid = 1e4, 2.0, 0
    trace, termination = [], 0
    g = (J * scale).T @ r
    if np.abs(J.T @ r)[mask].max() <= 1e-10:
        return P, trace, 1
    xn = x_norm(P, mask)
    it = 0
    while True:
        if it >= max_iter:
            termination = 0
            break
        if radius <= 1e-32:
            termination = 1
            break
        it += 1
        Js = (J * scale)[:, mask]
        H = Js.T @ Js
        gs = Js.T @ r
        D2 = np.clip(np.diag(H), 1e-6, 1e32) / radius
        try:
            step_s = np.linalg.solve(H + np.diag(D2), -gs)
            ok = bool(np.isfinite(step_s).all())
        except np.linalg.LinAlgError:
            ok = False
        model_change = -(step_s @ gs) - 0.5 * step_s @ H @ step_s if ok else 0.0
        if not ok or not model_change > 0:
            invalid += 1
            if invalid >= 5:
                termination = 2
                break
            radius *= 0.5
            trace.append((False, radius, cost))
            continue
        invalid = 0
        delta = np.zeros(J.shape[1])
        delta[mask] = step_s * scale[mask]
        cand = apply_step(olib, P, delta)
        cand_cost = dense_system(olib, cand, want_jac=False)[2]
        step_norm = math.sqrt(sum(float(((cand[k] - P[k]) ** 2).sum()) for k in ("pose", "mix")) + float(((cand["ext"] - P["ext"]) ** 2).sum())
                              + float(((cand["invdepth"] - P["invdepth"])[mask[15 * P["K"] + 7:]] ** 2).sum()))
        if step_norm <= 1e-8 * (xn + 1e-8):  # ParameterToleranceReached
            termination = 1
            trace.append((False, radius, cost))
            break
        if abs(cost - cand_cost) <= 1e-6 * cost:  # FunctionToleranceReached
            termination = 1
            trace.append((False, radius, cost))
            break
        rho = (cost - cand_cost) / model_change
        if rho > 1e-3:
            P = cand
            r, J, cost = dense_system(olib, P)
            xn = x_norm(P, mask)
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            decrease = 2.0
            trace.append((True, radius, cost))
            if np.abs(J.T @ r)[mask].max() <= 1e-10:
                termination = 1
                break
        else:
            radius /= decrease
            decrease *= 2.0
            trace.append((False, radius, cost))
    return P, trace, termination


def oracle_trace(olib, prob, max_iter):
    """Per-iteration (successful, radius, cost) of the oracle: re-solve from the same start with max_num_iterations = 1 .. max_iter."""
    trace, prev_succ, last = [], 0, None
    for k in range(1, max_iter + 1):
        q = copy.deepcopy(prob)
        s = oa.ba_solve(olib, q, k)
        if s["iterations"] < k:  # terminated earlier: no further iterations
            break
        trace.append((s["num_successful_steps"] > prev_succ, s["final_radius"], s["final_cost"]))
        prev_succ = s["num_successful_steps"]
        last = (q, s)
    return trace, last


def compare(olib, prob, max_iter):
    Pd, tr_d, term_d = dense_lm(olib, prob, max_iter)
    tr_o, _ = oracle_trace(olib, prob, max_iter)
    po = copy.deepcopy(prob)
    so = oa.ba_solve(olib, po, max_iter)
    assert so["iterations"] == len(tr_d), (so, len(tr_d))
    assert so["termination"] == (1 if term_d == 1 else 2 if term_d == 2 else 0)
    n = min(len(tr_o), len(tr_d))
    assert n >= len(tr_d) - 1  # a solve that stops on a tolerance reports that last iteration through its summary only
    for k in range(n):
        assert tr_o[k][0] == tr_d[k][0], f"iteration {k + 1}: accepted / rejected differs"
        assert abs(tr_o[k][1] - tr_d[k][1]) <= 1e-6 * tr_d[k][1], f"iteration {k + 1}: radius {tr_o[k][1]} vs {tr_d[k][1]}"
        assert abs(tr_o[k][2] - tr_d[k][2]) <= 1e-7 * tr_d[k][2], f"iteration {k + 1}: cost {tr_o[k][2]} vs {tr_d[k][2]}"
    assert so["num_successful_steps"] == sum(1 for t in tr_d if t[0])
    assert abs(so["final_cost"] - tr_d[-1][2]) <= 1e-7 * tr_d[-1][2]
    for key in ("pose", "invdepth", "ext"):
        assert np.abs(po[key] - Pd[key]).max() <= 1e-6 * np.abs(Pd[key]).max(), key
    mo, md = po["mix"].reshape(-1, 9), Pd["mix"].reshape(-1, 9)
    for sl in (slice(0, 3), slice(3, 6), slice(6, 9)):
        assert np.abs(mo[:, sl] - md[:, sl]).max() <= 1e-6 * np.abs(md[:, sl]).max()
    return Pd, po, tr_d


def test_full_two_pass_trajectory_cfg3(olib):
    """cfg 3 (K = 10, L = 300, free extrinsic + td), the reference's protocol: 5 iterations, chi2 culling, 15 iterations."""
    prob = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=10, L=300, seed=2024)[0]
    fc = prob["f_const"].reshape(-1, 14)
    fc[10, 3] += 0.2  # gross outliers so that the robust branch and both chi-square gates act
    fc[500, 4] -= 0.15
    prob["gnss_blh"][3:6] += np.array([1.0, -0.8, 0.5])
    prob["gnss_huber"] = 1
    Pd, po, tr1 = compare(olib, prob, 5)
    assert len(tr1) == 5
    # chi-square pass on BOTH results (IG/ic_gvins.cc:1241-1297); they must agree on what is removed / re-weighted
    outs = []
    for P in (Pd, po):
        rc, gc = oa.ba_residual_costs(olib, P)
        std = P["gnss_std"].reshape(-1, 3)
        rew = [g for g in range(P["n_gnss"]) if 2 * gc[g] > 7.815]
        for g in rew:
            std[g] *= math.sqrt(2 * gc[g] / 7.815)
        P["gnss_std"] = std.reshape(-1)
        out = 2 * rc > 5.991
        P["f_active"][out] = 0
        P["gnss_huber"] = 0
        outs.append((rew, np.nonzero(out)[0].tolist()))
    assert outs[0] == outs[1] and 10 in outs[0][1] and 500 in outs[0][1] and len(outs[0][0]) >= 1
    compare(olib, po, 15)


def test_trajectory_with_priors_and_rejections(olib):
    """First-window priors + marginalization prior + an initial guess far enough out that LM rejects steps (radius shrinks)."""
    prob = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=6, L=60, seed=41, with_priors=True, with_marg=True, pixel_noise=1.5)[0]
    rng = np.random.default_rng(3)
    prob["invdepth"] *= 1.0 + rng.normal(0, 0.6, prob["L"]).clip(-0.8, 3.0)
    prob["pose"].reshape(-1, 7)[:, :3] += rng.normal(0, 1.0, (prob["K"], 3))
    _, _, tr = compare(olib, prob, 20)
    assert any(not t[0] for t in tr), "the case must contain a rejected step"
