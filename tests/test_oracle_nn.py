"""CPU tests of the nonlinear-path oracle (oracle/nn_oracle.py): pinned to the decoded reference FNN fixture, to the layer
equations of the reference's NL modelers, and cross-checked three ways (forward-mode vs finite differences vs adjoint; twin
SQP vs an independent L-BFGS-B solve with a KKT certificate)."""
import dataclasses
import json
import pathlib

import numpy as np
import pytest

from conftest import load_nn_fixture
from oracle import mpc_oracle as mo
from oracle import nn_oracle as no

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _scenario(qt, n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(qt["xmin"], qt["xmax"], (n, 4)), rng.uniform(0.4, 1.0, (n, 4)), np.tile(qt["u_ref"], (n, 1))


def _design(m, qt, H):
    _, A, B = no.jacobian(m, qt["x_ref"][None], qt["u_ref"][None])
    P = mo.dare(A[0], B[0], qt["Q"], qt["R"])
    c = mo.condense(A[0], B[0], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"])
    return A[0], B[0], P, mo.auto_rho(c.Pc), c


def test_fnn_fixture_is_the_reference_layout(fnn_model):
    """fnn.jl:88-107: W_in 13x6 (no bias), one hidden (W 13x13, b 13), W_out 4x13; Float32 values promoted exactly."""
    g = json.loads((ROOT / "tests" / "golden" / "qt_fnn_model.json").read_text())
    m = fnn_model
    assert (m.nx, m.nu, m.n_neur, m.n_hid, m.activation, m.arch) == (4, 2, 13, 1, "relu", "fnn")
    for a in (m.W_in, m.W_h[0], m.b_h[0], m.W_out):
        assert np.array_equal(a, a.astype(np.float32).astype(np.float64))
    assert g["offsets"] == ["0x19ebd7", "0x19ed25", "0x19efd0", "0x19f010"]
    # the decoded chain is a one-step model of the quadruple tank: it nearly fixes the tests' operating point
    f = no.forward(m, np.full((1, 4), 0.65), np.full((1, 2), 1.2))[0]
    assert np.abs(f - 0.65).max() < 0.012
    assert no.reference_nl_variable_count(m, 15) == 3 * 4 * 16 + 3 * 2 * 15 + 13 * 2 * 15      # fnn.jl:111-119 at the tests' horizon 15


def test_layer_equations_match_reference_modelers(fnn_model, resnet_model):
    """Spell the reference's per-neuron constraints (fnn.jl:126-143, resnet.jl:125-142) out with Python loops."""
    rng = np.random.default_rng(3)
    for m in (fnn_model, resnet_model, dataclasses.replace(resnet_model, arch="polynet")):
        x, u = rng.uniform(0.2, 1.3, 4), rng.uniform(0, 3, 2)
        xu = np.concatenate([x, u])
        y = np.zeros((m.n_neur, m.n_hid + 1)); branch = np.zeros((m.n_neur, m.n_hid))
        for i in range(m.n_neur):
            y[i, 0] = m.W_in[i, :] @ xu
        for j in range(1, m.n_hid + 1):
            for i in range(m.n_neur):
                a = max(m.W_h[j - 1][i, :] @ y[:, j - 1] + m.b_h[j - 1][i], 0.0)
                branch[i, j - 1] = a
                y[i, j] = (y[i, j - 1] + a) if m.arch == "resnet" else a
            if m.arch == "polynet":                                   # polynet.jl:132-149, second path through the same W_j, b_j
                for i in range(m.n_neur):
                    y[i, j] = y[i, j - 1] + branch[i, j - 1] + max(m.W_h[j - 1][i, :] @ branch[:, j - 1] + m.b_h[j - 1][i], 0.0)
        assert np.allclose(m.W_out @ y[:, -1], no.forward(m, x[None], u[None])[0], rtol=0, atol=1e-14)


@pytest.mark.parametrize("activation", ["relu", "tanh", "sigmoid", "swish", "identity"])
def test_jacobian_forward_mode_vs_finite_differences(fnn_model, resnet_model, activation):
    rng = np.random.default_rng(5)
    for base in (fnn_model, resnet_model, dataclasses.replace(resnet_model, arch="polynet")):
        m = dataclasses.replace(base, activation=activation)
        x, u = rng.uniform(0.2, 1.3, (8, 4)), rng.uniform(0, 3, (8, 2))
        f, A, B = no.jacobian(m, x, u)
        assert np.allclose(f, no.forward(m, x, u))
        h = 1e-6
        for i in range(4):
            d = np.zeros(4); d[i] = h
            assert np.allclose((no.forward(m, x + d, u) - no.forward(m, x - d, u)) / (2 * h), A[:, :, i], atol=2e-7)
        for i in range(2):
            d = np.zeros(2); d[i] = h
            assert np.allclose((no.forward(m, x, u + d) - no.forward(m, x, u - d)) / (2 * h), B[:, :, i], atol=2e-7)


def test_densenet_layer_equations_and_jacobian():
    """densenet.jl:128-162 spelled out with loops (the new block is prepended; W_j is n x ((j-1) n), W_out nx x ((n_hid+1) n))."""
    rng = np.random.default_rng(8)
    n, nh, nx, nu = 5, 3, 3, 2
    m = no.NeuralModel("densenet", "swish", 0.4 * rng.standard_normal((n, nx + nu)), [0.3 * rng.standard_normal((n, (l + 1) * n)) for l in range(nh)],
                       [0.1 * rng.standard_normal(n) for _ in range(nh)], 0.3 * rng.standard_normal((nx, (nh + 1) * n)))
    x, u = rng.standard_normal(nx), rng.standard_normal(nu)
    y = [m.W_in @ np.concatenate([x, u])]
    for j in range(nh):
        a = np.array([no.act("swish", m.W_h[j][i, :] @ y[-1] + m.b_h[j][i]) for i in range(n)])
        y.append(np.concatenate([a, y[-1]]))
    assert np.allclose(m.W_out @ y[-1], no.forward(m, x[None], u[None])[0], rtol=0, atol=1e-14)
    f, A, B = no.jacobian(m, x[None], u[None]); h = 1e-6
    for i in range(nx):
        d = np.zeros(nx); d[i] = h
        assert np.allclose((no.forward(m, (x + d)[None], u[None]) - no.forward(m, (x - d)[None], u[None]))[0] / (2 * h), A[0, :, i], atol=1e-7)


def test_gradient_forward_sensitivities_equal_adjoint(qt, fnn_model):
    m = dataclasses.replace(fnn_model, activation="tanh")
    H, n = 12, 16
    _, _, P, _, _ = _design(m, qt, H)
    x0, xref, uref = _scenario(qt, n)
    rng = np.random.default_rng(1)
    u = rng.uniform(qt["umin"], qt["umax"], (n, H, 2))
    Hc = no.constant_hessian(2, H, qt["R"], 3.0 * np.eye(2))
    J1, g1, Pc, x, _Gall = no.linearize_trajectory(m, qt["Q"], P, Hc, u, x0, xref, uref)
    J2, g2 = no.grad_adjoint(m, qt["Q"], P, Hc, u, x0, xref, uref)
    J3, x3 = no.objective(m, qt["Q"], P, Hc, u, x0, xref, uref)
    assert np.allclose(J1, J2, rtol=1e-13) and np.allclose(J1, J3, rtol=1e-13) and np.allclose(x, x3)
    assert np.allclose(g1, g2, rtol=1e-10, atol=1e-9)
    assert np.abs(Pc - Pc.transpose(0, 2, 1)).max() < 1e-9 and np.linalg.eigvalsh(Pc).min() > 0
    assert no.reference_nl_residual(m, x, u) == 0.0


@pytest.mark.parametrize("fixture", ["qt_resnet_model.json", "qt_fnn_tanh_model.json", "qt_resnet_swish_model.json"])
def test_twin_sqp_matches_independent_solve(qt, fixture):
    """Acceptance ladder for the NL path: the twin's answer is a certified KKT point and equals an independent L-BFGS-B solve
    of the same NLP to the north-star tolerances (u0 within 1e-4 relative, objective within 1e-6 relative)."""
    m = load_nn_fixture(fixture)
    H, n = 20, 24
    _, _, P, rho, _ = _design(m, qt, H)
    x0, xref, uref = _scenario(qt, n)
    r = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0, xref, uref, rho)
    assert (r["status"] == 1).all() and r["iters"].max() <= 16
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"])
    _, g = no.grad_adjoint(m, qt["Q"], P, Hc, r["u"], x0, xref, uref)
    lb, ub = np.tile(qt["umin"], H), np.tile(qt["umax"], H)
    assert no.kkt_residual(g, r["u"].reshape(n, -1), lb, ub, tol=1e-7).max() < 5e-5 * np.abs(g).max()
    for i in range(6):
        u, J, k = no.nmpc_local_opt(m, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0[i], xref[i], uref[i], u_init=r["u"][i])
        assert k < 1e-5
        assert mo.u0_metric(r["u"][i, 0], u[0], qt["umin"], qt["umax"]) < 1e-4
        assert abs(J - r["objective"][i]) <= 1e-6 * abs(J)
        uc, Jc, _ = no.nmpc_local_opt(m, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0[i], xref[i], uref[i])     # cold start
        assert Jc >= r["objective"][i] * (1 - 1e-6) - 1e-9       # nothing better is found from the reference input either


def test_relu_fnn_fixture_kinks_are_flagged_not_hidden(qt, fnn_model):
    """The decoded relu FNN has kinks inside the operating box, so the NLP is piecewise quadratic and NONSMOOTH (the reason
    the reference also ships MILP modelers; its own NL-vs-MILP comparisons are `broken = true`,
    test/computation_mpc_test.jl:153-169).  Most problems end at a certified smooth KKT point that an independent solver
    confirms; the rest stop at a kink with the 'inaccurate' / max-iter status -- flagged, never reported as solved -- and
    still improve on the initial guess."""
    H, n = 20, 64
    _, _, P, rho, _ = _design(fnn_model, qt, H)
    x0, xref, uref = _scenario(qt, n)
    r = no.nmpc_sqp(fnn_model, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0, xref, uref, rho)
    ok = r["status"] == 1
    assert ok.mean() > 0.8 and set(np.unique(r["status"])) <= {1, 2, -2}
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"])
    J_init, _ = no.objective(fnn_model, qt["Q"], P, Hc, np.tile(uref[:, None, :], (1, H, 1)), x0, xref, uref)
    assert (r["objective"] <= J_init + 1e-9).all()
    for i in np.flatnonzero(ok)[:5]:
        u, J, k = no.nmpc_local_opt(fnn_model, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0[i], xref[i], uref[i], u_init=r["u"][i])
        assert abs(J - r["objective"][i]) <= 1e-6 * abs(J) and mo.u0_metric(r["u"][i, 0], u[0], qt["umin"], qt["umax"]) < 1e-4


def test_reference_relation_linear_vs_nonlinear(qt, fnn_model):
    """The one numeric statement the reference's tests make about this path (test/computation_mpc_test.jl:152,163): for the
    FNN model at horizon 5 from x0 = 0.6, the linear-method and NL-method state predictions agree within 0.5."""
    H = 5
    A, B, P, rho, c = _design(fnn_model, qt, H)
    p = mo.pack_params(qt["x0"], qt["x_ref"], qt["u_ref"])
    v, _ = mo.qp_exact(c, p[0])
    lin = mo.recover(c, v[None], p)
    nl = no.nmpc_sqp(fnn_model, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], qt["x0"][None], qt["x_ref"][None], qt["u_ref"][None], rho)
    assert np.abs(lin["x"] - nl["x"]).max() < 0.5 and np.abs(lin["e_x"] - nl["e_x"]).max() < 0.5


def test_twin_terminal_equality_vs_slsqp(qt):
    """NL model + terminal equality: the twin's solved problems satisfy e_H = 0 and match an independent SLSQP solve; the
    problems it flags infeasible cannot reach the reference at all (a pure feasibility search confirms)."""
    from scipy.optimize import minimize
    m = load_nn_fixture("qt_resnet_model.json")
    H, n = 20, 32
    _, _, P, rho, _ = _design(m, qt, H)
    rng = np.random.default_rng(4)
    xref = np.tile(qt["x_ref"], (n, 1)); x0 = xref + 0.03 * rng.standard_normal((n, 4)); uref = np.tile(qt["u_ref"], (n, 1))
    r = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], x0, xref, uref, rho, terminal="equality")
    ok = r["status"] == 1; bad = r["status"] == -3
    assert ok.sum() >= 8 and bad.sum() >= 1 and (ok | bad).all()
    assert np.abs(r["e_x"][ok][:, H]).max() < 1e-9
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"]); lb, ub = np.tile(qt["umin"], H), np.tile(qt["umax"], H)
    for i in np.flatnonzero(ok)[:3]:
        fg = lambda v: tuple(a[0] for a in no.grad_adjoint(m, qt["Q"], P, Hc, v.reshape(1, H, 2), x0[i:i + 1], xref[i:i + 1], uref[i:i + 1]))
        cons = lambda v: no.rollout(m, x0[i:i + 1], v.reshape(1, H, 2))[0, H] - xref[i]
        s_ = minimize(lambda v: (float(fg(v)[0]), fg(v)[1]), r["u"][i].ravel(), jac=True, method="SLSQP", bounds=list(zip(lb, ub)),
                      constraints=[{"type": "eq", "fun": cons}], options={"maxiter": 500, "ftol": 1e-15})
        assert np.abs(s_.x[:2] - r["u"][i, 0]).max() < 1e-4 * 4 and abs(s_.fun - r["objective"][i]) <= 1e-6 * abs(s_.fun)
    for i in np.flatnonzero(bad)[:2]:
        cons = lambda v: no.rollout(m, x0[i:i + 1], v.reshape(1, H, 2))[0, H] - xref[i]
        f_ = minimize(lambda v: float((cons(v) ** 2).sum()), np.tile(qt["u_ref"], H), method="L-BFGS-B", bounds=list(zip(lb, ub)),
                      options={"maxiter": 2000, "ftol": 1e-20, "gtol": 1e-14})
        assert np.sqrt(f_.fun) > 1e-5


def test_twin_contractive_terminal_set_vs_slsqp(qt):
    """NL model + mpc_terminal_ingredient = "contractive" (design_mpc.jl:333-340: e_H' e_H <= 0.9 e_0' e_0 added to the NL modeler's
    problem): the twin projects the linearised terminal rows onto the ball.  Sluggish weights (Q = I, R = 10 I, H = 3) make
    the constraint active on a good share of the problems; checked against an independent SLSQP solve of the same NLP."""
    from scipy.optimize import minimize
    m = load_nn_fixture("qt_fnn_tanh_model.json")
    H, n = 3, 48
    Q, R, S = np.eye(4), 10.0 * np.eye(2), np.zeros((2, 2))
    _, A, B = no.jacobian(m, qt["x_ref"][None], qt["u_ref"][None])
    P = mo.dare(A[0], B[0], Q, R)
    rho = mo.auto_rho(mo.condense(A[0], B[0], Q, R, S, P, H, qt["umin"], qt["umax"]).Pc)
    rng = np.random.default_rng(4)
    xref = np.tile(qt["x_ref"], (n, 1)); x0 = xref + rng.uniform(-0.15, 0.15, (n, 4)); uref = np.tile(qt["u_ref"], (n, 1))
    r = no.nmpc_sqp(m, Q, R, S, P, H, qt["umin"], qt["umax"], x0, xref, uref, rho, terminal="contractive")
    free = no.nmpc_sqp(m, Q, R, S, P, H, qt["umin"], qt["umax"], x0, xref, uref, rho)
    e0 = ((x0 - xref) ** 2).sum(1)
    ratio = (r["e_x"][:, H] ** 2).sum(1) / e0; ratio_free = (free["e_x"][:, H] ** 2).sum(1) / e0
    assert (r["status"] == 1).all() and (ratio <= 0.9 + 1e-8).all()
    viol = ratio_free > 0.9 + 1e-6
    assert 0.2 < viol.mean() < 0.8 and np.abs(ratio[viol] - 0.9).max() < 1e-7              # active exactly where the free optimum violates it
    assert np.abs(r["u"][~viol] - free["u"][~viol]).max() < 1e-5                             # and inactive elsewhere
    Hc = no.constant_hessian(2, H, R, S); lb, ub = np.tile(qt["umin"], H), np.tile(qt["umax"], H)
    for i in list(np.flatnonzero(viol)[:3]) + list(np.flatnonzero(~viol)[:1]):
        fg = lambda v: tuple(a[0] for a in no.grad_adjoint(m, Q, P, Hc, v.reshape(1, H, 2), x0[i:i + 1], xref[i:i + 1], uref[i:i + 1]))
        cons = lambda v: 0.9 * e0[i] - ((no.rollout(m, x0[i:i + 1], v.reshape(1, H, 2))[0, H] - xref[i]) ** 2).sum()
        s_ = minimize(lambda v: (float(fg(v)[0]), fg(v)[1]), np.tile(qt["u_ref"], H), jac=True, method="SLSQP", bounds=list(zip(lb, ub)),
                      constraints=[{"type": "ineq", "fun": cons}], options={"maxiter": 500, "ftol": 1e-15})
        assert mo.u0_metric(r["u"][i, 0], s_.x[:2], qt["umin"], qt["umax"]) < 1e-4 and abs(s_.fun - r["objective"][i]) <= 1e-6 * abs(s_.fun)


def test_relinearized_linear_method_is_the_reference_qp_per_problem(qt):
    """oracle.relinearized_linear_mpc (checker of mpcb_solve_relinearized_batch): per problem the linear modeler's QP on the
    Jacobian linearisation at THAT problem's reference with P = are(A_i, B_i, Q, R) (design_mpc.jl:319-327).  Checked
    against the reference's own sparse encoding of that QP (build_reference_qp) solved by the OSQP port."""
    from oracle import osqp_ref as orf
    import scipy.linalg as sla
    m = load_nn_fixture("qt_fnn_tanh_model.json")
    rng = np.random.default_rng(3)
    n, H = 4, 6
    x0 = rng.uniform(0.45, 0.95, (n, 4)); xref = rng.uniform(0.45, 0.9, (n, 4)); uref = rng.uniform(0.8, 2.2, (n, 2))
    orc = no.relinearized_linear_mpc(m, qt["Q"], qt["R"], qt["S"], H, qt["umin"], qt["umax"], x0, xref, uref)
    assert orc["solved"].all()
    for i in range(n):
        Ps = sla.solve_discrete_are(orc["A"][i], orc["B"][i], qt["Q"], qt["R"])
        assert np.abs(orc["P"][i] - Ps).max() < 1e-8 * np.abs(Ps).max()
        qp = mo.build_reference_qp(orc["A"][i], orc["B"][i], qt["Q"], qt["R"], qt["S"], orc["P"][i], H, xref[i], uref[i], x0[i], qt["umin"], qt["umax"])
        sol = orf.Workspace(orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u), orf.default_settings(eps_abs=1e-9, eps_rel=1e-9, max_iter=200000)).solve(cold_start=True)
        assert np.abs(sol["x"][qp.idx["u"]].T - orc["u"][i]).max() < 1e-5
        assert abs(sol["obj"] - orc["objective"][i]) <= 1e-6 * abs(orc["objective"][i])
