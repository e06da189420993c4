"""Known-answer tests for the marginalization restatement in the oracle (oracle/ba_ref.cpp: marginalize).  The reference ships
no tests and Eigen is absent (parity unpinned), so the oracle is pinned by: its Jacobi eigensolver against
numpy.linalg.eigh; the Schur complement against an independent numpy assembly from the oracle's single-factor evaluations;
the defining property of a marginal (min over the removed variables of the linearised cost == prior cost + const); and the
MarginalizationFactor round trip (feeding the prior back into the solver).  CPU only."""
import copy

import numpy as np
import pytest

from datagen import synth_ba
from tests import oracle_api as oa


@pytest.fixture(scope="module")
def olib(oracle):
    oa.declare_ba(oracle)
    return oracle


def make(olib, **kw):
    return synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), **kw)[0]


def test_jacobi_eig_vs_numpy(olib):
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 40, 97):
        B = rng.normal(size=(n, n + 3)) * np.exp(rng.uniform(-6, 6, size=(n, 1)))
        A = B @ B.T
        ev, V = oa.sym_eig(olib, A)
        ref = np.linalg.eigvalsh(A)
        assert np.allclose(np.sort(ev), ref, rtol=1e-10, atol=1e-12 * ref.max())
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
        assert np.abs(A @ V - V * ev[None, :]).max() <= 1e-12 * np.abs(A).max() * n


def numpy_equation(olib, prob, num_marg):
    """H0, b0 of MarginalizationInfo::constructEquation assembled in numpy from single-factor oracle evaluations, in the column
    order the oracle documents: marginalized [pose_k, mix_k (k < num_marg), landmarks asc] then [pose_k, mix_k ..., ext, td]."""
    K = prob["K"]
    pose, mix, ext = prob["pose"].reshape(-1, 7), prob["mix"].reshape(-1, 9), prob["ext"]
    act = prob["f_active"].astype(bool) & (prob["f_ref"] < num_marg)
    lms = sorted(set(prob["f_lm"][act].tolist()))
    touched = set()
    for f in np.nonzero(act)[0]:
        touched.add(("pose", int(prob["f_obs"][f])))
    if num_marg - 1 < prob["n_imu"]:
        touched.add(("pose", num_marg)); touched.add(("mix", num_marg))
    gs = {0: "pose", 1: "mix"}
    if prob["marg_r"] > 0:
        for t, nd in zip(prob["marg_block_type"], prob["marg_block_node"]):
            if t < 2 and nd >= num_marg:
                touched.add((gs[int(t)], int(nd)))
    col = {}
    idx = 0
    for k in range(num_marg):
        col[("pose", k)] = idx; idx += 6
        col[("mix", k)] = idx; idx += 9
    for l in lms:
        col[("lm", l)] = idx; idx += 1
    m = idx
    for k in range(num_marg, K):
        for nm, sz in (("pose", 6), ("mix", 9)):
            if (nm, k) in touched:
                col[(nm, k)] = idx; idx += sz
    col[("ext", 0)] = idx; idx += 6
    col[("td", 0)] = idx; idx += 1
    n0 = idx
    H = np.zeros((n0, n0)); b = np.zeros(n0)

    def add(blocks, Js, r):
        for (ka, Ja) in zip(blocks, Js):
            for (kb, Jb) in zip(blocks, Js):
                H[col[ka]:col[ka] + Ja.shape[1], col[kb]:col[kb] + Jb.shape[1]] += Ja.T @ Jb
            b[col[ka]:col[ka] + Ja.shape[1]] -= Ja.T @ r

    off, pn = prob["pn_off"], prob["pn"].reshape(-1, 4)
    for k in range(min(num_marg, prob["n_imu"])):
        r, Js = oa.imu_eval(olib, prob["imu_blob"].reshape(-1, 480)[k], pn[off[k]:off[k + 1]], pose[k], mix[k], pose[k + 1], mix[k + 1])
        add([("pose", k), ("mix", k), ("pose", k + 1), ("mix", k + 1)], [Js[0][:, :6], Js[1], Js[2][:, :6], Js[3]], r)
    for g, nd in enumerate(prob["gnss_node"]):
        if nd < num_marg:
            r = np.zeros(3); J = np.zeros((3, 7))
            a = [pose[nd].copy(), prob["gnss_blh"][3 * g:3 * g + 3].copy(), prob["gnss_std"][3 * g:3 * g + 3].copy(), np.array(prob["lever"], np.float64)]
            olib.icgo_gnss_eval(oa._p(a[0]), oa._p(a[1]), oa._p(a[2]), oa._p(a[3]), oa._p(r), oa._p(J))
            add([("pose", int(nd))], [J[:, :6]], r)
    for f in np.nonzero(act)[0]:
        i, j, l = int(prob["f_ref"][f]), int(prob["f_obs"][f]), int(prob["f_lm"][f])
        r, Js = oa.reproj_eval(olib, pose[i], pose[j], ext[:7], prob["invdepth"][l], ext[7], prob["f_const"][14 * f:14 * f + 14], prob["reproj_std"])
        add([("pose", i), ("pose", j), ("ext", 0), ("lm", l), ("td", 0)], [Js[0][:, :6], Js[1][:, :6], Js[2][:, :6], Js[3], Js[4]], r)
    if prob["marg_r"] > 0:
        rr = prob["marg_r"]
        J0 = prob["marg_J0"].reshape(rr, rr); e0 = prob["marg_e0"]
        blocks, Js, dx = [], [], np.zeros(rr)
        c = xo = 0
        for t, nd in zip(prob["marg_block_type"], prob["marg_block_node"]):
            t = int(t); nd = int(nd)
            ls = {0: 6, 1: 9, 2: 6, 3: 1}[t]; g = {0: 7, 1: 9, 2: 7, 3: 1}[t]
            x0 = prob["marg_x0"][xo:xo + g]
            x = pose[nd] if t == 0 else mix[nd] if t == 1 else ext[:7] if t == 2 else ext[7:8]
            if g == 7:
                dx[c:c + 3] = x[:3] - x0[:3]
                q0 = np.array([-x0[3], -x0[4], -x0[5], x0[6]]) / np.dot(x0[3:7], x0[3:7])
                dq = synth_ba.q_mul(q0, x[3:7])
                dx[c + 3:c + 6] = 2 * dq[:3] * (1 if dq[3] >= 0 else -1)
            else:
                dx[c:c + ls] = x - x0
            blocks.append((["pose", "mix", "ext", "td"][t], nd if t < 2 else 0)); Js.append(J0[:, c:c + ls])
            c += ls; xo += g
        add(blocks, Js, e0 + J0 @ dx)
    if prob["has_pose_prior"] or prob["has_mix_prior"]:
        raise NotImplementedError
    return H, b, m, col


@pytest.mark.parametrize("with_marg", [False, True])
def test_schur_complement_vs_numpy(olib, with_marg):
    prob = make(olib, K=10, L=300, seed=11, with_marg=with_marg)
    out = oa.ba_marginalize(olib, prob, 1)
    H, b, m, col = numpy_equation(olib, prob, 1)
    assert out["m"] == m and out["r"] == H.shape[0] - m
    Hmm = 0.5 * (H[:m, :m] + H[:m, :m].T)
    ev, V = np.linalg.eigh(Hmm)
    inv = V @ np.diag(np.where(ev > 1e-8, 1.0 / np.where(ev > 1e-8, ev, 1.0), 0.0)) @ V.T
    Hp = H[m:, m:] - H[m:, :m] @ inv @ H[:m, m:]
    bp = b[m:] - H[m:, :m] @ inv @ b[:m]
    sc = np.sqrt(np.abs(np.diag(Hp)))
    sc[sc == 0] = 1
    # Hrr - Hrm Hmm^-1 Hmr cancels ~1e3..1e5 : 1 and cond(Hmm) ~ 1e8, so two correct eigensolvers agree to ~1e-7 (diag-scaled)
    assert np.abs((out["Hp"] - Hp) / np.outer(sc, sc)).max() < 2e-6
    assert np.abs((out["bp"] - bp) / sc).max() < 2e-6 * max(1.0, np.abs(bp / sc).max())
    # linearization: J0^T J0 == Hp on the retained spectrum, J0^T e0 == -bp projected on it
    J0, e0 = out["J0"], out["e0"]
    S, U = np.linalg.eigh(out["Hp"])
    keep = S > 1e-8
    Hp_k = (U[:, keep] * S[keep]) @ U[:, keep].T
    assert np.abs((J0.T @ J0 - Hp_k) / np.outer(sc, sc)).max() < 1e-9
    bp_k = U[:, keep] @ (U[:, keep].T @ out["bp"])
    assert np.abs((J0.T @ e0 + bp_k) / sc).max() < 1e-7 * max(1.0, np.abs(bp / sc).max())


def test_prior_is_the_marginal_of_the_linearised_cost(olib):
    """min over dx_m of 0.5 dx^T H dx - b^T dx  ==  0.5 |e0 + J0 dx_r|^2 + const, for any dx_r"""
    prob = make(olib, K=6, L=60, seed=5)
    out = oa.ba_marginalize(olib, prob, 1)
    H, b, m, _ = numpy_equation(olib, prob, 1)
    rng = np.random.default_rng(0)
    sc = 1.0 / np.sqrt(np.diag(H)[m:])

    def full_min(dxr):
        dxm = np.linalg.solve(H[:m, :m], b[:m] - H[:m, m:] @ dxr)
        dx = np.concatenate([dxm, dxr])
        return 0.5 * dx @ H @ dx - b @ dx

    def prior(dxr):
        e = out["e0"] + out["J0"] @ dxr
        return 0.5 * e @ e

    d0 = full_min(np.zeros(out["r"])) - prior(np.zeros(out["r"]))
    for _ in range(5):
        dxr = rng.normal(size=out["r"]) * sc
        d = full_min(dxr) - prior(dxr)
        assert abs(d - d0) <= 1e-7 * max(1.0, abs(prior(dxr)))


def test_block_list_and_x0(olib):
    prob = make(olib, K=10, L=300, seed=11, with_marg=True)
    out = oa.ba_marginalize(olib, prob, 1)
    # remained: every pose that observes a landmark anchored in node 0, pose_1 + mix_1 (IMU factor 0 and the old prior), ext, td
    types, nodes = out["block_type"].tolist(), out["block_node"].tolist()
    assert types[-2:] == [2, 3]
    assert (0, 0) in zip(types, nodes) and (1, 0) in zip(types, nodes)
    assert sum(1 for t in types if t == 1) == 1
    obs = set((prob["f_obs"][prob["f_ref"] == 0] - 1).tolist())
    assert obs <= set(n for t, n in zip(types, nodes) if t == 0)
    assert out["r"] == sum({0: 6, 1: 9, 2: 6, 3: 1}[t] for t in types)
    pose = prob["pose"].reshape(-1, 7)
    assert np.array_equal(out["x0"][:7], pose[1])


def test_prior_round_trip_through_the_solver(olib):
    """Solve a window, marginalize node 0, drop node 0 + its landmarks and re-solve with the prior: the remaining states must not
    move (the prior carries exactly the information that was removed, and the window was at its optimum)."""
    prob = make(olib, K=8, L=120, seed=9, pixel_noise=0.3)
    prob["reproj_huber"] = 0
    prob["gnss_huber"] = 0
    for _ in range(4):  # Ceres' function_tolerance stops each call early; repeat until the window sits at its optimum
        oa.ba_solve(olib, prob, 50)
    out = oa.ba_marginalize(olib, prob, 1)
    K = prob["K"]
    keep_f = prob["f_ref"] >= 1
    q = copy.deepcopy(prob)
    q.update(K=K - 1, pose=prob["pose"][7:].copy(), mix=prob["mix"][9:].copy(), F=int(keep_f.sum()),
             f_lm=prob["f_lm"][keep_f].copy(), f_ref=(prob["f_ref"][keep_f] - 1).astype(np.int32), f_obs=(prob["f_obs"][keep_f] - 1).astype(np.int32),
             f_const=prob["f_const"].reshape(-1, 14)[keep_f].reshape(-1).copy(), f_active=prob["f_active"][keep_f].copy(),
             n_imu=prob["n_imu"] - 1, imu_blob=prob["imu_blob"][480:].copy(),
             pn_off=(prob["pn_off"][1:] - prob["pn_off"][1]).astype(np.int32), pn=prob["pn"][4 * prob["pn_off"][1]:].copy())
    g = prob["gnss_node"] >= 1
    q.update(n_gnss=int(g.sum()), gnss_node=(prob["gnss_node"][g] - 1).astype(np.int32), gnss_blh=prob["gnss_blh"].reshape(-1, 3)[g].reshape(-1).copy(),
             gnss_std=prob["gnss_std"].reshape(-1, 3)[g].reshape(-1).copy())
    q.update(marg_r=out["r"], marg_nblocks=len(out["block_type"]), marg_block_type=out["block_type"], marg_block_node=out["block_node"],
             marg_x0=out["x0"], marg_J0=out["J0"].reshape(-1).copy(), marg_e0=out["e0"])
    before = q["pose"].copy()
    for _ in range(2):
        oa.ba_solve(olib, q, 20)
    assert np.abs(q["pose"].reshape(-1, 7)[:, :3] - before.reshape(-1, 7)[:, :3]).max() < 5e-4


def test_two_node_marginalization_is_a_marginal(olib):
    """num_marg = 2 (a non-keyframe time node sits between the two oldest keyframes, ic_gvins.cc:1421-1424): both nodes, both
    preintegration factors and the landmarks anchored in either node leave; the marginal property must still hold"""
    prob = make(olib, K=7, L=70, seed=21)
    out = oa.ba_marginalize(olib, prob, 2)
    H, b, m, col = numpy_equation(olib, prob, 2)
    assert out["m"] == m and out["r"] == H.shape[0] - m
    assert ("mix", 2) in col and ("pose", 2) in col  # node 2 is touched by preintegration factor 1
    assert not any(t == 1 and n > 0 for t, n in zip(out["block_type"], out["block_node"]))  # only mix of the new node 0 remains
    rng = np.random.default_rng(1)
    sc = 1.0 / np.sqrt(np.diag(H)[m:])

    def full_min(dxr):
        dxm = np.linalg.solve(H[:m, :m], b[:m] - H[:m, m:] @ dxr)
        dx = np.concatenate([dxm, dxr])
        return 0.5 * dx @ H @ dx - b @ dx

    def prior(dxr):
        e = out["e0"] + out["J0"] @ dxr
        return 0.5 * e @ e

    d0 = full_min(np.zeros(out["r"])) - prior(np.zeros(out["r"]))
    for _ in range(4):
        dxr = rng.normal(size=out["r"]) * sc
        assert abs(full_min(dxr) - prior(dxr) - d0) <= 1e-7 * max(1.0, abs(prior(dxr)))
