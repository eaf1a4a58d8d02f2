"""CPU tests: pin the oracle against everything the reference gives us (structure, fixture, known answers) and prove
the two encodings of the problem (reference sparse model vs condensed QP) have the same optimum."""
import json
import pathlib

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import mpc_oracle as mo
from oracle import osqp_ref as orf

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_fixture_bits_roundtrip():
    g = json.loads((ROOT / "tests" / "golden" / "qt_linear_model.json").read_text())
    A = np.array(g["A"]); B = np.array(g["B"])
    Ah = np.array([[np.frombuffer(bytes.fromhex(h), "<f4")[0] for h in row] for row in g["A_f32_hex"]], np.float64)
    Bh = np.array([[np.frombuffer(bytes.fromhex(h), "<f4")[0] for h in row] for row in g["B_f32_hex"]], np.float64)
    assert np.array_equal(A, Ah) and np.array_equal(B, Bh)          # Float32 -> Float64 promotion is exact
    assert abs(A[0, 0] - 0.968072) < 1e-6 and abs(B[3, 0] - 0.0144097) < 1e-7     # SURVEY 8c decoded values
    ev = np.linalg.eigvals(A)
    assert np.abs(ev).max() < 1.0 and abs(np.abs(ev).max() - 0.9771) < 2e-3


def test_dare_three_ways(qt, mpc):
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    Ps = sla.solve_discrete_are(qt["A"], qt["B"], qt["Q"], qt["R"])
    Pl = mpc.dare(qt["A"], qt["B"], qt["Q"], qt["R"])       # host code inside libmpcb200 (no GPU needed)
    assert np.abs(P - Ps).max() < 1e-7 and np.abs(Pl - Ps).max() < 1e-8
    # SURVEY 8c(3) known answers
    assert abs(Ps[0, 0] - 1547.2440077) < 1e-6 and abs(Ps[0, 1] + 233.25836909) < 1e-7 and abs(Ps[3, 3] - 515.53206304) < 1e-7
    rng = np.random.default_rng(1)
    G = rng.standard_normal((12, 12)); A = 0.95 * G / np.abs(np.linalg.eigvals(G)).max(); B = rng.standard_normal((12, 3)) / 4
    Pl = mpc.dare(A, B, 100 * np.eye(12), 0.1 * np.eye(3)); Ps = sla.solve_discrete_are(A, B, 100 * np.eye(12), 0.1 * np.eye(3))
    assert np.abs(Pl - Ps).max() < 1e-8 * np.abs(Ps).max()


@pytest.mark.parametrize("terminal,count", [("none", 74), ("contractive", 75), ("equality", 78)])
def test_reference_constraint_counts(qt, terminal, count):
    """test/terminal_ingredient_test.jl:160,237,317 at H=5, nx=4, nu=2."""
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, 5, qt["x_ref"], qt["u_ref"], qt["x0"], qt["umin"], qt["umax"],
                               terminal=terminal)
    assert qp.n_jump_constraints == count
    # variable containers and shapes (test/modeler_implementation_test.jl:86-105): 6 containers, x/e_x/x_reference 4x6, u/e_u/u_reference 2x5
    assert len(qp.idx) == 6 and qp.idx["x"].shape == (4, 6) and qp.idx["e_u"].shape == (2, 5)
    assert qp.P.shape[0] == 3 * 4 * 6 + 3 * 2 * 5
    if terminal == "equality":
        # the 4 terminal rows sit at positions 55..58 of the AffExpr-EqualTo group (terminal_ingredient_test.jl:318-321)
        assert (qp.rows["terminal"].start + 1, qp.rows["terminal"].stop) == (55, 58)


def test_reference_counts_with_state_constraint_and_S(qt):
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], 2.0 * np.eye(2), P, 5, qt["x_ref"], qt["u_ref"], qt["x0"], qt["umin"], qt["umax"],
                               qt["xmin"], qt["xmax"], state_constraint=True)
    assert qp.n_jump_constraints == 74 + 2 * 4 * 6 + 2 * 4          # + state bounds on all H+1 columns + delta_u rows (i < H)
    assert "delta_u" in qp.idx and len(qp.idx) == 7               # design_mpc.jl:423-427


@pytest.mark.parametrize("H", [5, 20, 50])
def test_kat_unconstrained_lqr(qt, H):
    """from x0 = 0.6 no bound is active, so u0* is the infinite-horizon LQR law for every H."""
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
    p = mo.pack_params(qt["x0"], qt["x_ref"], qt["u_ref"])
    v, info = mo.qp_exact(c, p[0])
    assert info["n_active"] == 0
    assert np.allclose(v[:2], [2.75594127, 2.95507466], atol=1e-7)
    K = np.linalg.solve(qt["R"] + qt["B"].T @ P @ qt["B"], qt["B"].T @ P @ qt["A"])
    assert np.allclose(v[:2], qt["u_ref"] - K @ (qt["x0"] - qt["x_ref"]), atol=1e-9)


def _solve_sparse(qp, eps=1e-9, max_iter=200000):
    prob = orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u)
    w = orf.Workspace(prob, orf.default_settings(eps_abs=eps, eps_rel=eps, max_iter=max_iter, check_termination=25))
    return w.solve(cold_start=True)


@pytest.mark.parametrize("terminal,Sval,state_c", [("none", 0.0, False), ("equality", 0.0, False), ("none", 3.0, False), ("none", 0.0, True)])
def test_sparse_reference_model_equals_condensed(qt, terminal, Sval, state_c):
    """(i) == (ii): the reference's redundant multiple-shooting encoding and the condensed QP the GPU solves share the optimum."""
    H = 6
    S = Sval * np.eye(2)
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    rng = np.random.default_rng(7)
    for trial in range(3):
        xref = rng.uniform(0.5, 0.9, 4)
        x0 = xref + (0.0003 if terminal == "equality" else 0.15) * rng.standard_normal(4)
        x0 = np.clip(x0, qt["xmin"] + 0.01, qt["xmax"] - 0.01)
        qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], S, P, H, xref, qt["u_ref"], x0, qt["umin"], qt["umax"], qt["xmin"], qt["xmax"],
                                   state_constraint=state_c, terminal=terminal)
        r = _solve_sparse(qp)
        assert r["status"] == 1
        u_sparse = r["x"][qp.idx["u"].T.ravel()]
        c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], S, P, H, qt["umin"], qt["umax"], qt["xmin"], qt["xmax"], state_constraint=state_c,
                        terminal=terminal)
        p = mo.pack_params(x0, xref, qt["u_ref"])
        if state_c:     # qp_exact has no inequality general rows: use the condensed ADMM twin at tight tolerance
            v = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-9, eps_rel=1e-9, check_every=10, max_iter=200000))["v"][0]
        else:
            v, _ = mo.qp_exact(c, p[0])
        tol = 3e-5 if terminal == "equality" else 2e-6      # the terminal-equality QP is ill-conditioned in u (|B| ~ 1e-2)
        assert np.abs(u_sparse - v).max() < tol, (terminal, Sval, state_c, np.abs(u_sparse - v).max())
        rec = mo.recover(c, v[None], p)
        assert np.abs(r["x"][qp.idx["x"].T.ravel()] - rec["x"].ravel()).max() < 2e-6
        assert abs(r["obj"] - rec["objective"][0]) <= 1e-5 * max(1.0, abs(rec["objective"][0]))     # same J, constants included


def test_osqp_port_kkt_solve_matches_dense(qt):
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, 5, qt["x_ref"], qt["u_ref"], qt["x0"], qt["umin"], qt["umax"])
    st = orf.default_settings(scaling=0)
    for ordering in ("mindeg", "natural"):
        prob = orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u, ordering=ordering)
        w = orf.Workspace(prob, st)
        n, m = prob.n, prob.m
        rho = np.where(np.abs(qp.u - qp.l) < 1e-4, 1e3 * 0.1, np.where((qp.l < -1e26) & (qp.u > 1e26), 1e-6, 0.1))
        K = np.block([[qp.P.toarray() + 1e-6 * np.eye(n), qp.A.toarray().T], [qp.A.toarray(), -np.diag(1.0 / rho)]])
        b = np.random.default_rng(0).standard_normal(n + m)
        assert np.abs(w.kkt_solve(b) - np.linalg.solve(K, b)).max() < 1e-7 * np.abs(np.linalg.solve(K, b)).max()
    assert orf.Workspace(orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u), st).nnz_L() <= orf.Workspace(orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u, ordering="natural"), st).nnz_L()


def test_osqp_port_default_accuracy_ladder(qt):
    """Oracle ladder (iii): OSQP at its defaults (what the reference actually runs) lands within its own tolerance of the
    exact optimum -- and no closer: this is the accuracy of the reference path."""
    H, n = 20, 64
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["x_ref"], qt["u_ref"], qt["x0"], qt["umin"], qt["umax"])
    prob = orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u)
    rng = np.random.default_rng(0)
    x0 = rng.uniform(qt["xmin"], qt["xmax"], (n, 4)); xref = rng.uniform(0.4, 1.0, (n, 4))
    rows = np.concatenate([qp.x0_rows, qp.xref_rows]); vals = np.hstack([x0, np.tile(xref, (1, H + 1))])
    sel = qp.idx["u"].T.ravel()
    r = orf.solve_batch(prob, orf.default_settings(), rows, vals, sel, cold_start=True, nthreads=2)
    assert (r["status"] == 1).all() and r["iters"].max() <= 4000
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
    p = mo.pack_params(x0, xref, qt["u_ref"])
    ex = np.array([mo.qp_exact(c, p[i])[0] for i in range(n)])
    err = mo.u0_metric(r["x"][:, :2], ex[:, :2], qt["umin"], qt["umax"])
    assert err.max() < 0.1 and np.median(err) < 5e-3
    tight = orf.solve_batch(prob, orf.default_settings(eps_abs=1e-7, eps_rel=1e-7, max_iter=100000), rows, vals, sel, cold_start=True, nthreads=2)
    assert mo.u0_metric(tight["x"][:, :2], ex[:, :2], qt["umin"], qt["umax"]).max() < 1e-4


def test_osqp_port_flags_infeasible(qt):
    H = 5
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["x_ref"], qt["u_ref"], [1.3, 1.3, 0.25, 0.25], qt["umin"], qt["umax"],
                               terminal="equality")
    r = _solve_sparse(qp, eps=1e-3, max_iter=4000)
    assert r["status"] in (-3, 3)


def test_twin_matches_exact_and_warm_start_helps(qt):
    H, n = 20, 128
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
    rng = np.random.default_rng(2)
    x0 = rng.uniform(qt["xmin"], qt["xmax"], (n, 4)); xref = rng.uniform(0.4, 1.0, (n, 4))
    p = mo.pack_params(x0, xref, qt["u_ref"])
    s = mo.AdmmSettings(eps_abs=1e-7, eps_rel=1e-7, check_every=5)
    tw = mo.admm_condensed(c, p, s)
    assert (tw["status"] == 1).all()
    ex = np.array([mo.qp_exact(c, p[i], v_init=tw["v"][i])[0] for i in range(n)])
    assert mo.u0_metric(tw["v"][:, :2], ex[:, :2], qt["umin"], qt["umax"]).max() < 1e-4
    J = mo.recover(c, tw["v"], p)["objective"]; Jex = mo.recover(c, ex, p)["objective"]
    assert (np.abs(J - Jex) / np.abs(Jex)).max() < 1e-6
    warm = mo.admm_condensed(c, p, s, v0=tw["v"], y0=tw["y"])
    assert warm["iters"].max() <= s.check_every and (warm["status"] == 1).all()


def test_cold_start_point_saves_iterations_and_keeps_the_solution(qt):
    """settings.cold_init = 1 (opt-in): a cold start at the clipped unconstrained optimum with the one-number dual guess needs fewer iterations
    than OSQP's zeros and ends at the same optimum (both within the parity tolerance of the exact solve); where no bound is active it IS
    the optimum and the first check terminates."""
    import dataclasses
    H, n = 20, 256
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
    rng = np.random.default_rng(4)
    x0 = rng.uniform(qt["xmin"], qt["xmax"], (n, 4)); xref = rng.uniform(0.4, 1.0, (n, 4))
    p = mo.pack_params(x0, xref, qt["u_ref"])
    s0 = mo.AdmmSettings(eps_abs=1e-7, eps_rel=1e-7, check_every=5, sigma=0.0)
    assert s0.cold_init == 0                          # opt-in: the default stays OSQP's cold start
    s1 = dataclasses.replace(s0, cold_init=1)
    r1, r0 = mo.admm_condensed(c, p, s1), mo.admm_condensed(c, p, s0)
    assert (r1["status"] == 1).all() and (r0["status"] == 1).all()
    assert r1["iters"].mean() < 0.95 * r0["iters"].mean() and r1["iters"].max() <= r0["iters"].max()
    ex = np.array([mo.qp_exact(c, p[i], v_init=r1["v"][i])[0] for i in range(64)])
    for r in (r0, r1): assert mo.u0_metric(r["v"][:64, :2], ex[:, :2], qt["umin"], qt["umax"]).max() < 1e-4
    assert np.abs(r1["v"] - r0["v"]).max() < 5e-5
    # the starting point itself: inside the box the dual guess is zero, on the box it has the multiplier's sign
    x, y = mo.cold_start_point(c, p, r1["rho"], s1.init_kappa)
    vunc = -np.linalg.solve(c.Pc, (p @ c.Lq.T).T).T
    assert np.abs(x - np.clip(vunc, c.lb, c.ub)).max() < 1e-9 and (y[(x > c.lb) & (x < c.ub)] == 0).all()
    assert (y[x >= c.ub] >= 0).all() and (y[x <= c.lb] <= 0).all()      # OSQP's sign: positive on an active upper bound
    # no active bound: one check
    wide = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, -1e3 * np.ones(2), 1e3 * np.ones(2))
    rw = mo.admm_condensed(wide, p, s1)
    assert (rw["iters"] == s1.check_every).all() and np.abs(rw["v"] - vunc).max() < 1e-9


def test_contractive_ball_twin_vs_slsqp(qt):
    """Terminal "contractive" (design_mpc.jl:333-340) in the condensed twin: the ball projection inside the ADMM reproduces an
    independent SLSQP solve of the QCQP, and the constraint is active on part of the batch."""
    from scipy.optimize import minimize
    H, n = 3, 96
    Q, R = np.eye(4), 10 * np.eye(2)
    P = mo.dare(qt["A"], qt["B"], Q, R)
    c = mo.condense(qt["A"], qt["B"], Q, R, qt["S"], P, H, qt["umin"], qt["umax"], terminal="contractive")
    assert c.nball == 4 and c.mg == 4
    rng = np.random.default_rng(3)
    xref = rng.uniform(0.5, 0.9, (n, 4)); x0 = xref + 0.15 * rng.standard_normal((n, 4))
    p = mo.pack_params(x0, xref, qt["u_ref"])
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-8, eps_rel=1e-8, check_every=5, max_iter=20000))
    assert (tw["status"] == 1).all()
    rec = mo.recover(c, tw["v"], p)
    e0 = np.linalg.norm(rec["e_x"][:, 0], axis=1); eH = np.linalg.norm(rec["e_x"][:, H], axis=1)
    active = np.abs(eH - np.sqrt(0.9) * e0) < 1e-6
    assert (eH <= np.sqrt(0.9) * e0 + 1e-7).all() and active.sum() >= 8
    for i in np.flatnonzero(active)[:3]:
        q = c.Lq @ p[i]; b = c.Lb @ p[i]; r2 = 0.9 * e0[i] ** 2
        r = minimize(lambda v: (0.5 * v @ c.Pc @ v + q @ v, c.Pc @ v + q), tw["v"][i], jac=True, method="SLSQP", bounds=list(zip(c.lb, c.ub)),
                     constraints=[{"type": "ineq", "fun": lambda v: r2 - np.sum((c.G @ v - b) ** 2), "jac": lambda v: -2 * (c.G @ v - b) @ c.G}],
                     options={"maxiter": 1000, "ftol": 1e-16})
        assert np.abs(r.x - tw["v"][i]).max() < 1e-5


@pytest.mark.parametrize("H,sigma", [(5, 0.0), (20, 1e-6), (60, 0.0)])
def test_stagewise_riccati_x_update_equals_condensed_operator(qt, H, sigma):
    """The stage-wise (Riccati) x-update of csrc/admm_riccati.cu is the SAME linear map as the condensed operator
    T = (Pc + (sigma + rho) I)^-1 of the other kernels: one backward and one forward sweep reproduce T r to round-off."""
    P = mo.dare(qt["A"], qt["B"], qt["Q"], qt["R"])
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
    s = mo.AdmmSettings(sigma=sigma)
    T, _, _, rho = mo.admm_matrices(c, s)
    fac = mo.riccati_factors(c, sigma, rho)
    r = np.random.default_rng(H).standard_normal((64, c.nz))
    t = mo.riccati_apply(c, fac, r)
    assert np.abs(t - r @ T).max() <= 1e-12 * np.abs(t).max()
