"""GPU parity tests of the nonlinear path: CUDA network rollout / Jacobians and the batched SQP kernel (through the C ABI)
against oracle/nn_oracle.py on the same seeded inputs."""
import dataclasses

import numpy as np
import pytest

from conftest import load_nn_fixture
from oracle import mpc_oracle as mo
from oracle import nn_oracle as no

pytestmark = pytest.mark.gpu

U0_TOL, OBJ_TOL = 1e-4, 1e-6        # north_star tolerances


def to_chain(mpc, m):
    cls = {"fnn": mpc.Fnn, "resnet": mpc.ResNet, "polynet": mpc.PolyNet, "densenet": mpc.DenseNet}[m.arch]
    return cls(m.W_in, list(zip(m.W_h, m.b_h)), m.W_out, activation=m.activation)


def make_system(mpc, qt, m):
    return mpc.ConstrainedBlackBoxControlDiscreteSystem(to_chain(mpc, m), 4, 2, mpc.Hyperrectangle(qt["xmin"], qt["xmax"]),
                                                        mpc.Hyperrectangle(qt["umin"], qt["umax"]))


def scenario(qt, n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(qt["xmin"], qt["xmax"], (n, 4)), rng.uniform(0.4, 1.0, (n, 4)), qt["u_ref"].copy()


@pytest.mark.parametrize("activation", ["relu", "tanh", "sigmoid", "swish", "identity"])
def test_rollout_and_jacobian_match_oracle(mpc, fnn_model, resnet_model, activation):
    rng = np.random.default_rng(7)
    for base in (fnn_model, resnet_model, dataclasses.replace(resnet_model, arch="polynet")):
        m = dataclasses.replace(base, activation=activation)
        f = to_chain(mpc, m)
        n, H = 517, 9                                      # ragged: not a multiple of the CTA's 4 warps
        x0 = rng.uniform(0.2, 1.3, (n, 4)); u = rng.uniform(0, 4, (n, H, 2))
        x = f.rollout(x0, u)
        xo = no.rollout(m, x0, u)
        assert np.abs(x - xo).max() < 1e-13 * max(1.0, np.abs(xo).max())
        fx, A, B = f.jacobian(x0, u[:, 0])
        fo, Ao, Bo = no.jacobian(m, x0, u[:, 0])
        assert np.abs(fx - fo).max() < 1e-13 and np.abs(A - Ao).max() < 1e-13 and np.abs(B - Bo).max() < 1e-13
        assert np.allclose(f(np.concatenate([x0[0], u[0, 0]])), fo[0], atol=1e-13)      # system.f([x; u])


def test_deeper_wider_network(mpc):
    """Generic sizes: nx = 3, nu = 2, 40 neurons (more than one warp), 3 hidden layers."""
    rng = np.random.default_rng(11)
    for arch in ("fnn", "resnet", "polynet", "densenet"):
        dense = arch == "densenet"
        m = no.NeuralModel(arch, "tanh", 0.3 * rng.standard_normal((40, 5)), [0.2 * rng.standard_normal((40, 40 * (l + 1 if dense else 1))) for l in range(3)],
                           [0.1 * rng.standard_normal(40) for _ in range(3)], 0.2 * rng.standard_normal((3, 40 * (4 if dense else 1))))
        f = to_chain(mpc, m)
        x0 = rng.standard_normal((33, 3)); u = rng.standard_normal((33, 6, 2))
        assert np.abs(f.rollout(x0, u) - no.rollout(m, x0, u)).max() < 1e-12
        _, A, B = f.jacobian(x0, u[:, 0]); _, Ao, Bo = no.jacobian(m, x0, u[:, 0])
        assert np.abs(A - Ao).max() < 1e-13 and np.abs(B - Bo).max() < 1e-13


def test_linear_method_on_blackbox_model(mpc, qt, fnn_model):
    """fnn.jl:37-55: LinearProgramming on a black-box model = Jacobian linearisation at the reference + the linear modeler;
    P from the linearisation at the last reference column (design_mpc.jl:312-327)."""
    sys_ = make_system(mpc, qt, fnn_model)
    C = mpc.proceed_controller(sys_, "model_predictive_control", 5, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="linear", mpc_b200_eps_abs=1e-8, mpc_b200_eps_rel=1e-8, mpc_b200_check_every=5)
    _, A, B = no.jacobian(fnn_model, qt["x_ref"][None], qt["u_ref"][None])
    P = mo.dare(A[0], B[0], qt["Q"], qt["R"])
    assert np.abs(C.tuning.terminal_ingredient.P - P).max() < 1e-6 * np.abs(P).max()
    d = C.tuning.modeler.design()
    c = mo.condense(A[0], B[0], qt["Q"], qt["R"], qt["S"], C.tuning.terminal_ingredient.P, 5, qt["umin"], qt["umax"])
    assert np.allclose(d["Pc"], c.Pc, rtol=1e-10, atol=1e-9)
    mpc.update_initialization(C, qt["x0"]); mpc.calculate(C)
    v, _ = mo.qp_exact(c, mo.pack_params(qt["x0"], qt["x_ref"], qt["u_ref"])[0])
    assert mo.u0_metric(C.computation_results.u[:, 0], v[:2], qt["umin"], qt["umax"]) < U0_TOL


@pytest.mark.parametrize("fixture,H", [("qt_resnet_model.json", 20), ("qt_fnn_tanh_model.json", 20), ("qt_resnet_swish_model.json", 20), ("qt_fnn_tanh_model.json", 7),
                                       ("qt_resnet_swish_model.json", 33), ("qt_polynet_tanh_model.json", 20),
                                       ("qt_densenet_tanh_model.json", 20)])
def test_sqp_matches_twin_and_independent_solve(mpc, qt, fixture, H):
    m = load_nn_fixture(fixture)
    n = 300
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear")
    mod = C.tuning.modeler
    x0, xref, uref = scenario(qt, n)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    # design parity
    d = mod.design()
    _, A, B = no.jacobian(m, qt["x_ref"][None], qt["u_ref"][None])
    P = mo.dare(A[0], B[0], qt["Q"], qt["R"])
    assert np.abs(d["A"] - A[0]).max() < 1e-13 and np.abs(d["B"] - B[0]).max() < 1e-13 and np.abs(d["P"] - P).max() < 1e-6 * np.abs(P).max()
    # twin: same algorithm, same settings
    tw = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), d["rho"])
    assert (res["status"] == tw["status"]).mean() > 0.98 and set(np.unique(res["status"])) <= {1, -2}
    ok = (res["status"] == 1) & (tw["status"] == 1)
    assert ok.all() if ("polynet" not in fixture and "densenet" not in fixture) else ok.mean() > 0.9     # slow GN convergence on a few problems (cap of 20)
    assert (res["iters"] == tw["iters"]).mean() > 0.9
    assert np.abs(res["u"][ok] - tw["u"][ok]).max() < 5e-6            # both stop at ||step|| <= 1e-6 of the same fixed point
    assert np.abs(res["objective"][ok] - tw["objective"][ok]).max() <= 1e-9 * np.abs(tw["objective"]).max()
    # outputs are consistent with the reference's variables: x = rollout(u), e_x = x - x_ref, e_u = u - u_ref
    assert np.abs(res["x"] - no.rollout(m, x0, res["u"])).max() < 1e-12 and no.reference_nl_residual(m, res["x"], res["u"]) < 1e-12
    assert np.abs(res["e_x"] - (res["x"] - xref[:, None, :])).max() < 1e-15 and np.abs(res["e_u"] - (res["u"] - uref)).max() < 1e-15
    assert np.abs(res["u0"] - res["u"][:, 0]).max() == 0.0 and np.abs(C.computation_results.u - res["u"][0].T).max() == 0.0
    assert (res["u"] >= qt["umin"] - 2e-9).all() and (res["u"] <= qt["umax"] + 2e-9).all()       # x~ of the last QP: within eps_abs of the box
    # independent solve (Ipopt stand-in) + KKT certificate
    for i in np.flatnonzero(ok)[:5]:
        u, J, k = no.nmpc_local_opt(m, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0[i], xref[i], uref, u_init=res["u"][i])
        assert k < 1e-5
        assert mo.u0_metric(res["u0"][i], u[0], qt["umin"], qt["umax"]) < U0_TOL
        assert abs(J - res["objective"][i]) <= OBJ_TOL * abs(J)


def test_relu_fnn_fixture_statuses_and_reference_relation(mpc, qt, fnn_model):
    """The reference's own fixture (relu FNN): nonsmooth NLP.  Statuses equal the twin's; solved problems match it; and the
    one numeric relation the reference's tests state holds (|x_linear - x_nl| <= 0.5 at H = 5 from x0 = 0.6,
    test/computation_mpc_test.jl:152,163)."""
    sys_ = make_system(mpc, qt, fnn_model)
    H, n = 20, 256
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_programming_type="non_linear")
    x0, xref, uref = scenario(qt, n)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    d = C.tuning.modeler.design()
    tw = no.nmpc_sqp(fnn_model, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), d["rho"])
    assert set(np.unique(res["status"])) <= {1, 2, -2} and (res["status"] == 1).mean() > 0.8
    same = (res["status"] == tw["status"]) & (res["iters"] == tw["iters"])
    assert same.mean() > 0.9
    ok = same & (res["status"] == 1)
    assert np.abs(res["u"][ok] - tw["u"][ok]).max() < 5e-6
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"])
    J_init, _ = no.objective(fnn_model, qt["Q"], d["P"], Hc, np.tile(uref, (n, H, 1)), x0, xref, np.tile(uref, (n, 1)))
    assert (res["objective"] <= J_init + 1e-9).all()
    # reference relation at H = 5
    Cl = mpc.proceed_controller(sys_, "model_predictive_control", 5, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_programming_type="linear")
    Cn = mpc.proceed_controller(sys_, "model_predictive_control", 5, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_programming_type="non_linear")
    for Cx in (Cl, Cn):
        mpc.update_initialization(Cx, qt["x0"]); mpc.calculate(Cx)
    assert np.abs(Cl.computation_results.x - Cn.computation_results.x).max() < 0.5
    assert np.abs(Cl.computation_results.e_x - Cn.computation_results.e_x).max() < 0.5


def test_sqp_warm_start_and_ragged_batches(mpc, qt, resnet_model):
    m = load_nn_fixture("qt_fnn_tanh_model.json")
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", 20, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear")
    mod = C.tuning.modeler
    x0, xref, uref = scenario(qt, 131)
    full = mod.solve_batch(x0, xref, uref, want=("u", "u0", "objective", "y"))
    for nb in (1, 3, 130):
        part = mod.solve_batch(x0[:nb], xref[:nb], uref, want=("u", "objective"))
        assert np.array_equal(part["u"], full["u"][:nb]) and np.array_equal(part["iters"], full["iters"][:nb])     # per-problem results do not depend on the batch
    warm = mod.solve_batch(x0, xref, uref, want=("u", "objective"), warm=(full["u"], full["y"]))
    assert (warm["iters"] == 1).all() and (warm["status"] == 1).all() and np.abs(warm["u"] - full["u"]).max() < 2e-6
    assert warm["inner_iters"].mean() < 0.2 * full["inner_iters"].mean()
    # unsupported configurations are refused, not approximated
    for kw in ({"mpc_terminal_ingredient": "neighborhood"},):
        with pytest.raises(mpc.MpcbError):
            mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", 20, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                                   mpc_programming_type="non_linear", **kw)


@pytest.mark.parametrize("fixture", ["qt_resnet_model.json", "qt_fnn_tanh_model.json"])
def test_sqp_terminal_equality(mpc, qt, fixture):
    """mpc_terminal_ingredient = "equality" on an NL model (design_mpc.jl:330-331: e_x[:,end] == 0 added to the NL modeler's
    problem).  CUDA vs twin; solved problems reach the reference exactly and match an independent SLSQP solve of the same
    NLP; problems whose terminal state is unreachable under the input box are flagged -3 by both."""
    from scipy.optimize import minimize
    m = load_nn_fixture(fixture)
    H, n = 20, 96
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear", mpc_terminal_ingredient="equality")
    mod = C.tuning.modeler
    rng = np.random.default_rng(4)
    xref = np.tile(qt["x_ref"], (n, 1)); x0 = xref + 0.02 * rng.standard_normal((n, 4)); uref = qt["u_ref"].copy()
    res = mod.solve_batch(x0, xref, uref, want=("u", "u0", "x", "e_x", "objective", "y"))
    d = mod.design()
    tw = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), d["rho"], terminal="equality")
    assert set(np.unique(res["status"])) <= {1, -3, 2, -2}
    assert (res["status"] == tw["status"]).mean() > 0.95
    ok = (res["status"] == 1) & (tw["status"] == 1)
    assert ok.sum() >= n // 3 and (res["status"] == -3).sum() == (tw["status"] == -3).sum()
    assert np.abs(res["e_x"][ok][:, H]).max() < 1e-9                       # the terminal state IS the reference
    assert np.abs(res["u"][ok] - tw["u"][ok]).max() < 5e-6 and np.abs(res["objective"][ok] - tw["objective"][ok]).max() <= 1e-8 * np.abs(tw["objective"][ok]).max()
    assert np.abs(res["y"][ok][:, 2 * H:] - tw["y_terminal"][ok]).max() <= 1e-4 * max(1.0, np.abs(tw["y_terminal"][ok]).max())
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"])
    lb, ub = np.tile(qt["umin"], H), np.tile(qt["umax"], H)
    for i in np.flatnonzero(ok)[:4]:
        fg = lambda v: tuple(a[0] for a in no.grad_adjoint(m, qt["Q"], d["P"], Hc, v.reshape(1, H, 2), x0[i:i + 1], xref[i:i + 1], uref[None]))
        cons = lambda v: no.rollout(m, x0[i:i + 1], v.reshape(1, H, 2))[0, H] - xref[i]
        r = minimize(lambda v: (float(fg(v)[0]), fg(v)[1]), res["u"][i].ravel(), jac=True, method="SLSQP", bounds=list(zip(lb, ub)),
                     constraints=[{"type": "eq", "fun": cons}], options={"maxiter": 500, "ftol": 1e-15})
        assert np.abs(cons(r.x)).max() < 1e-9
        assert mo.u0_metric(res["u0"][i], r.x[:2], qt["umin"], qt["umax"]) < U0_TOL and abs(r.fun - res["objective"][i]) <= OBJ_TOL * abs(r.fun)



@pytest.mark.parametrize("fixture,method", [("qt_fnn_tanh_model.json", "non_linear"), ("qt_resnet_swish_model.json", "non_linear"), ("qt_fnn_tanh_model.json", "linear")])
def test_sqp_contractive_terminal_set(mpc, qt, fixture, method):
    """mpc_terminal_ingredient = "contractive" on an NL model (design_mpc.jl:333-340 added to the NL modeler's problem -- the
    reference hands this NLP to Ipopt): the SQP kernel projects the linearised terminal rows onto the ball.  Sluggish weights
    (Q = I, R = 10 I, H = 3) make the set active.  CUDA vs twin, and vs an independent SLSQP solve of the same NLP.
    method = "linear": the same constraint on the per-problem re-linearised controllers, against the linear path's oracle."""
    from scipy.optimize import minimize
    m = load_nn_fixture(fixture)
    H, n = 3, 200
    Q, R, S = np.eye(4), 10.0 * np.eye(2), np.zeros((2, 2))
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear", mpc_terminal_ingredient="contractive", mpc_Q=1.0, mpc_R=10.0)
    mod = C.tuning.modeler
    rng = np.random.default_rng(4)
    xref = np.tile(qt["x_ref"], (n, 1)); x0 = xref + rng.uniform(-0.15, 0.15, (n, 4)); uref = qt["u_ref"].copy()
    e0 = ((x0 - xref) ** 2).sum(1)
    if method == "linear":
        res = mod.solve_batch(x0, xref, uref, want=("u", "u0", "e_x", "objective"), method="linear")
        assert (res["status"] == 1).all()
        _, A, B = no.jacobian(m, qt["x_ref"][None], qt["u_ref"][None]); P = mo.dare(A[0], B[0], Q, R)
        c = mo.condense(A[0], B[0], Q, R, S, P, H, qt["umin"], qt["umax"], terminal="contractive")
        p = mo.pack_params(x0, xref, uref)
        tw = mo.admm_condensed(c, p, mo.AdmmSettings(eps_abs=1e-10, eps_rel=1e-10, check_every=5, max_iter=50000))
        assert (tw["status"] == 1).all()
        rec = mo.recover(c, tw["v"], p)
        ratio = (res["e_x"][:, H] ** 2).sum(1) / e0
        assert (ratio <= 0.9 + 1e-7).all() and 0.1 < (ratio > 0.9 - 1e-6).mean() < 0.9
        assert mo.u0_metric(res["u0"], rec["u"][:, 0], qt["umin"], qt["umax"]).max() < U0_TOL
        assert (np.abs(res["objective"] - rec["objective"]) <= OBJ_TOL * np.abs(rec["objective"])).all()
        return
    res = mod.solve_batch(x0, xref, uref, want=("u", "u0", "x", "e_x", "objective", "y"))
    d = mod.design()
    tw = no.nmpc_sqp(m, Q, R, S, d["P"], H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), d["rho"], terminal="contractive")
    assert (res["status"] == tw["status"]).mean() > 0.98
    ok = (res["status"] == 1) & (tw["status"] == 1)
    assert ok.mean() > 0.95
    assert np.abs(res["u"][ok] - tw["u"][ok]).max() < 5e-6 and np.abs(res["objective"][ok] - tw["objective"][ok]).max() <= 1e-8 * np.abs(tw["objective"][ok]).max()
    ratio = (res["e_x"][:, H] ** 2).sum(1) / e0
    assert (ratio[ok] <= 0.9 + 1e-7).all()
    act = ok & (ratio > 0.9 - 1e-6)
    assert 0.2 < act.mean() < 0.9                                          # the set is active on a good share, inactive on the rest
    assert np.abs(res["y"][ok & ~act][:, 2 * H:]).max() < 1e-5             # multipliers of the terminal rows vanish where it is inactive
    Hc = no.constant_hessian(2, H, R, S); lb, ub = np.tile(qt["umin"], H), np.tile(qt["umax"], H)
    for i in list(np.flatnonzero(act)[:3]) + list(np.flatnonzero(ok & ~act)[:1]):
        fg = lambda v: tuple(a[0] for a in no.grad_adjoint(m, Q, d["P"], Hc, v.reshape(1, H, 2), x0[i:i + 1], xref[i:i + 1], uref[None]))
        cons = lambda v: 0.9 * e0[i] - ((no.rollout(m, x0[i:i + 1], v.reshape(1, H, 2))[0, H] - xref[i]) ** 2).sum()
        r = minimize(lambda v: (float(fg(v)[0]), fg(v)[1]), np.tile(uref, H), jac=True, method="SLSQP", bounds=list(zip(lb, ub)),
                     constraints=[{"type": "ineq", "fun": cons}], options={"maxiter": 500, "ftol": 1e-15})
        assert mo.u0_metric(res["u0"][i], r.x[:2], qt["umin"], qt["umax"]) < U0_TOL and abs(r.fun - res["objective"][i]) <= OBJ_TOL * abs(r.fun)


@pytest.mark.parametrize("fixture,terminal", [("qt_resnet_model.json", "none"), ("qt_fnn_tanh_model.json", "none"), ("qt_resnet_model.json", "equality")])
def test_sqp_state_constraint(mpc, qt, fixture, terminal):
    """kw `mpc_state_constraint` on an NL model (fnn.jl:146-154 / resnet.jl:145-153): linearised state-box rows inside the SQP.
    The box is tightened to [0.55, 0.75] with references partly beyond it so the rows are active.  CUDA vs twin, and vs an
    independent SLSQP solve of the same NLP with the nonlinear state constraints."""
    from scipy.optimize import minimize
    m = load_nn_fixture(fixture)
    H, n = 20, 96
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedBlackBoxControlDiscreteSystem(to_chain(mpc, m), 4, 2, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear", mpc_state_constraint=True, mpc_terminal_ingredient=terminal)
    mod = C.tuning.modeler
    rng = np.random.default_rng(5)
    if terminal == "none":
        xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    else:                                    # reachable terminal state: start near the design reference, upper bound just above it
        xref = np.tile(qt["x_ref"], (n, 1)); x0 = xref + 0.01 * rng.standard_normal((n, 4))
    uref = qt["u_ref"].copy()
    res = mod.solve_batch(x0, xref, uref, want=("u", "u0", "x", "e_x", "objective", "y"))
    d = mod.design()
    tw = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), d["rho"], terminal=terminal,
                     xmin=xmin, xmax=xmax, state_constraint=True)
    assert res["y"].shape == (n, 2 * H + 4 * H + (4 if terminal == "equality" else 0))
    assert (res["status"] == tw["status"]).mean() > 0.95
    ok = (res["status"] == 1) & (tw["status"] == 1)
    assert ok.mean() > (0.9 if terminal == "none" else 0.3)
    assert np.abs(res["u"][ok] - tw["u"][ok]).max() < 5e-6 and np.abs(res["objective"][ok] - tw["objective"][ok]).max() <= 1e-8 * np.abs(tw["objective"][ok]).max()
    xs = res["x"][ok]
    assert (xs[:, 1:] <= xmax + 1e-8).all() and (xs[:, 1:] >= xmin - 1e-8).all()
    if terminal == "none":
        touching = (xs[:, 1:] > xmax - 1e-6).any(axis=(1, 2))
        assert touching.sum() >= n // 2                                    # the rows are really active
        pick = np.flatnonzero(ok)[np.flatnonzero(touching)[:3]]
    else:
        assert np.abs(res["e_x"][ok][:, H]).max() < 1e-9
        pick = np.flatnonzero(ok)[:2]
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"]); lb, ub = np.tile(qt["umin"], H), np.tile(qt["umax"], H)
    for i in pick:
        fg = lambda v: tuple(a[0] for a in no.grad_adjoint(m, qt["Q"], d["P"], Hc, v.reshape(1, H, 2), x0[i:i + 1], xref[i:i + 1], uref[None]))
        roll = lambda v: no.rollout(m, x0[i:i + 1], v.reshape(1, H, 2))[0]
        cons = [{"type": "ineq", "fun": lambda v: np.concatenate([(xmax - roll(v)[1:]).ravel(), (roll(v)[1:] - xmin).ravel()])}]
        if terminal == "equality": cons.append({"type": "eq", "fun": lambda v: roll(v)[H] - xref[i]})
        r = minimize(lambda v: (float(fg(v)[0]), fg(v)[1]), res["u"][i].ravel(), jac=True, method="SLSQP", bounds=list(zip(lb, ub)), constraints=cons,
                     options={"maxiter": 500, "ftol": 1e-15})
        assert mo.u0_metric(res["u0"][i], r.x[:2], qt["umin"], qt["umax"]) < U0_TOL and abs(r.fun - res["objective"][i]) <= OBJ_TOL * abs(r.fun)


@pytest.mark.parametrize("fixture,kw", [("qt_fnn_tanh_model.json", {}), ("qt_resnet_model.json", {"mpc_state_constraint": True})])
def test_closed_loop_on_gpu_equals_host_loop(mpc, qt, fixture, kw):
    """SURVEY 8f-1 with the network as the plant: solve -> apply u[:,1] to system.f -> re-solve, kept on the device for a batch of
    plants, must reproduce the same loop driven from the host through the batched solve and rollout entries (same shifted warm
    start).  The regulated plants approach their references."""
    m = load_nn_fixture(fixture)
    H, n, steps = 10, 150, 8
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear", **kw)
    mod = C.tuning.modeler
    f = mod.nn
    rng = np.random.default_rng(2)
    x0 = rng.uniform(0.5, 0.8, (n, 4)); xref = np.tile(qt["x_ref"], (n, 1)) + rng.uniform(-0.03, 0.03, (n, 4)); uref = qt["u_ref"].copy()
    for warm in (True, False):
        dev = mod.closed_loop(x0, xref, uref, steps, warm_start=warm)
        x = x0.copy(); w = None
        xs, us, its, bad = [x.copy()], [], np.zeros(n, np.int64), np.zeros(n, np.int64)
        for t in range(steps):
            r = mod.solve_batch(x, xref, uref, want=("u", "u0", "y"), warm=w)
            its += r["inner_iters"]; bad += r["status"] != 1
            x = f.rollout(x, r["u0"][:, None, :])[:, 1]
            xs.append(x.copy()); us.append(r["u0"].copy())
            if warm:
                wu = np.concatenate([r["u"][:, 1:], r["u"][:, -1:]], 1)
                wy = r["y"].copy(); yb = wy[:, :2 * H].reshape(n, H, 2); wy[:, :2 * H] = np.concatenate([yb[:, 1:], yb[:, -1:]], 1).reshape(n, -1)
                w = (wu, wy)
        xs = np.stack(xs, 1); us = np.stack(us, 1)
        assert np.array_equal(dev["x_traj"], xs) and np.array_equal(dev["u_traj"], us)          # same kernels, same inputs: bit-identical
        assert np.array_equal(dev["iters_total"], its) and np.array_equal(dev["unsolved_steps"], bad)
        assert (bad == 0).mean() > 0.9
        if warm: it_warm = its.mean()
        else: assert it_warm < 0.95 * its.mean()                                # the shifted warm start pays
    good = bad == 0
    assert np.abs(dev["x_traj"][good, -1] - xref[good]).mean() < np.abs(x0 - xref)[good].mean()          # slow plant, 8 steps: the mean tracking error shrinks
