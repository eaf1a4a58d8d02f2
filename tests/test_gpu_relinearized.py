"""GPU parity tests of the device-side design for many systems: batched Riccati equations
(mpcb_dare_batch) and the per-problem re-linearised linear method on a black-box model (mpcb_solve_relinearized_batch),
through the C ABI, against oracle/ (per-problem Jacobian -> DARE -> condensed QP -> exact solve) on the same seeded inputs."""
import numpy as np
import pytest
import scipy.linalg as sla

from conftest import load_nn_fixture
from oracle import mpc_oracle as mo
from oracle import nn_oracle as no
from test_gpu_nmpc import make_system, to_chain

pytestmark = pytest.mark.gpu

U0_TOL, OBJ_TOL = 1e-4, 1e-6        # north_star tolerances


def references(qt, n, seed=3):
    rng = np.random.default_rng(seed)
    x0 = rng.uniform(0.45, 0.95, (n, 4))
    xref = rng.uniform(0.45, 0.9, (n, 4))
    uref = rng.uniform(0.8, 2.2, (n, 2))
    return x0, xref, uref


def test_dare_batch_matches_scipy_and_value_iteration(mpc, qt):
    rng = np.random.default_rng(0)
    # (a) the linearisations of a trained network at many references
    m = load_nn_fixture("qt_fnn_tanh_model.json")
    _, xref, uref = references(qt, 257)
    _, A, B = no.jacobian(m, xref, uref)
    P, st = mpc.dare_batch(A, B, qt["Q"], qt["R"])
    assert (st > 0).all() and st.max() <= 40
    for i in range(0, 257, 16):
        Ps = sla.solve_discrete_are(A[i], B[i], qt["Q"], qt["R"])
        assert np.abs(P[i] - Ps).max() < 1e-9 * np.abs(Ps).max()
    Pv = mo.dare(A[5], B[5], qt["Q"], qt["R"])                         # the oracle's own (value iteration)
    assert np.abs(P[5] - Pv).max() < 1e-8 * np.abs(Pv).max()
    assert np.abs(P - P.transpose(0, 2, 1)).max() == 0.0
    Ph = mpc.dare(A[7], B[7], qt["Q"], qt["R"])                        # the host's doubling solver: same recurrence
    assert np.abs(P[7] - Ph).max() < 1e-10 * np.abs(Ph).max()
    # (b) generic sizes incl. unstable open loops, nx not a power of two, nu > nx
    for nx, nu in ((7, 3), (3, 5), (12, 3), (1, 1)):
        n = 41
        A = rng.standard_normal((n, nx, nx)) * (1.2 / np.sqrt(nx)); B = rng.standard_normal((n, nx, nu))
        Q = np.eye(nx) * 3.0; R = np.diag(rng.uniform(0.1, 2.0, nu))
        P, st = mpc.dare_batch(A, B, Q, R)
        assert (st > 0).all()
        for i in range(0, n, 5):
            Ps = sla.solve_discrete_are(A[i], B[i], Q, R)
            assert np.abs(P[i] - Ps).max() < 1e-8 * np.abs(Ps).max()
    # (c) a system with an uncontrollable unstable mode has no stabilising solution: flagged per system, the others unaffected
    A = np.tile(np.diag([1.5, 0.5]), (3, 1, 1)); B = np.tile(np.array([[0.0], [1.0]]), (3, 1, 1))
    A[1] = np.diag([0.9, 0.5])
    P, st = mpc.dare_batch(A, B, np.eye(2), np.eye(1))
    assert st[0] == -1 and st[2] == -1 and st[1] > 0 and np.isnan(P[0]).all() and np.isfinite(P[1]).all()
    with pytest.raises(mpc.MpcbError, match="singular"):
        mpc.dare_batch(A, B, np.eye(2), np.zeros((1, 1)))


@pytest.mark.parametrize("fixture,H", [("qt_fnn_tanh_model.json", 10), ("qt_resnet_swish_model.json", 20), ("qt_densenet_tanh_model.json", 5)])
def test_relinearized_matches_per_problem_design(mpc, qt, fixture, H):
    m = load_nn_fixture(fixture)
    n = 150
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear")
    mod = C.tuning.modeler
    x0, xref, uref = references(qt, n)
    res = mod.solve_batch(x0, xref, uref, want=("u", "e_u", "x", "e_x", "u0", "objective", "y"), method="linear")
    assert (res["status"] == 1).all() and (res["iters"] == 1).all()
    orc = no.relinearized_linear_mpc(m, qt["Q"], qt["R"], qt["S"], H, qt["umin"], qt["umax"], x0, xref, uref)
    assert mo.u0_metric(res["u0"], orc["u"][:, 0], qt["umin"], qt["umax"]).max() < U0_TOL
    assert np.abs(res["u"] - orc["u"]).max() < 1e-6
    assert (np.abs(res["objective"] - orc["objective"]) <= OBJ_TOL * np.abs(orc["objective"])).all()
    assert np.abs(res["x"] - orc["x"]).max() < 1e-6 and np.abs(res["e_x"] - (res["x"] - xref[:, None, :])).max() < 1e-15
    assert np.abs(res["e_u"] - (res["u"] - uref[:, None, :])).max() < 1e-15
    assert (res["u"] >= qt["umin"] - 2e-9).all() and (res["u"] <= qt["umax"] + 2e-9).all()
    active = ((res["u"] <= qt["umin"] + 1e-7) | (res["u"] >= qt["umax"] - 1e-7)).any(axis=(1, 2))
    assert 0.2 < active.mean()                                          # the input box matters in this scenario
    # the same problem through the single-system linear path of this library, designed at that problem's reference
    for i in (0, 77):
        Cl = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(xref[i]), list(uref[i]), mpc_solver="b200",
                                    mpc_programming_type="linear", mpc_b200_eps_abs=1e-9, mpc_b200_eps_rel=1e-9)
        assert np.abs(Cl.tuning.terminal_ingredient.P - orc["P"][i]).max() < 1e-7 * np.abs(orc["P"][i]).max()
        mpc.update_initialization(Cl, x0[i]); mpc.calculate(Cl)
        assert np.abs(Cl.computation_results.u.T - res["u"][i]).max() < 1e-6
    # device-pointer entry on the caller's stream gives the same answer (ragged batch: not a multiple of the CTA width)
    import torch
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    k = n - 7
    dx0, dxr, dur = dev(x0[:k]), dev(xref[:k]), dev(uref[:k])
    du = torch.empty((k, H, 2), dtype=torch.float64, device="cuda"); dst = torch.empty(k, dtype=torch.int32, device="cuda"); dit = torch.empty_like(dst)
    io = mpc._lib.BatchIO()
    io.batch = k; io.x0 = dx0.data_ptr(); io.xref = dxr.data_ptr(); io.uref = dur.data_ptr(); io.u = du.data_ptr(); io.status = dst.data_ptr(); io.iters = dit.data_ptr()
    mod.solve_batch_device(io, stream=torch.cuda.current_stream().cuda_stream, method="linear")
    torch.cuda.synchronize()
    assert (dst.cpu().numpy() == 1).all() and np.abs(du.cpu().numpy() - res["u"][:k]).max() == 0.0


def test_relinearized_constraints(mpc, qt):
    """State box + terminal equality rows on the per-problem designs."""
    m = load_nn_fixture("qt_fnn_tanh_model.json")
    H, n = 10, 48
    rng = np.random.default_rng(9)
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedBlackBoxControlDiscreteSystem(to_chain(mpc, m), 4, 2, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    uref = rng.uniform(1.0, 2.0, (n, 2))
    for terminal, sc in (("none", True), ("equality", False)):
        if sc:          # references partly beyond the state box: the rows are active
            xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
        else:           # the terminal equality must be reachable inside the input box
            xref = rng.uniform(0.6, 0.7, (n, 4)); x0 = xref + rng.uniform(-0.004, 0.004, (n, 4))
        C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                                   mpc_programming_type="non_linear", mpc_state_constraint=sc, mpc_terminal_ingredient=terminal, mpc_b200_max_iter=20000)
        res = C.tuning.modeler.solve_batch(x0, xref, uref, want=("u", "u0", "x", "e_x", "objective"), method="linear")
        orc = no.relinearized_linear_mpc(m, qt["Q"], qt["R"], qt["S"], H, qt["umin"], qt["umax"], x0, xref, uref, xmin=xmin, xmax=xmax,
                                         state_constraint=sc, terminal=terminal)
        ok = (res["status"] == 1) & orc["solved"]
        assert ok.mean() > 0.9 and (res["status"][~orc["solved"]] != 1).all()
        assert mo.u0_metric(res["u0"][ok], orc["u"][ok][:, 0], qt["umin"], qt["umax"]).max() < U0_TOL
        assert (np.abs(res["objective"][ok] - orc["objective"][ok]) <= OBJ_TOL * np.abs(orc["objective"][ok])).all()
        if sc:
            assert (res["x"][ok][:, 1:] <= xmax + 1e-7).all() and (res["x"][ok][:, 1:] >= xmin - 1e-7).all()
            assert (res["x"][ok][:, 1:] > xmax - 1e-6).any(axis=(1, 2)).mean() > 0.5
        if terminal == "equality": assert np.abs(res["e_x"][ok][:, H]).max() < 1e-7


def test_relinearized_design_failure_is_reported_per_problem(mpc):
    """A model whose linearisation has an unstable mode the input cannot reach has no stabilising Riccati solution: the host
    design refuses it (mpcb_create_nmpc without P), and the per-problem device design reports it as data (status -20) while the
    call itself succeeds -- mirroring how the linear path reports solver outcomes (status[] of OSQP's codes)."""
    # identity activation: f(x, u) = [1.5 x1, 0.5 x2 + u]
    f = mpc.Fnn(np.eye(3), [(np.eye(3), np.zeros(3))], np.array([[1.5, 0.0, 0.0], [0.0, 0.5, 1.0]]), activation="identity")
    Q, R, S = np.eye(2), np.eye(1), np.zeros((1, 1))
    args = (Q, R, S)
    with pytest.raises(mpc.MpcbError, match="dare"):
        mpc.B200NonlinearModeler(f, *args, None, [-1.0], [1.0], None, None, 5, np.zeros(2), np.zeros(1))
    mod = mpc.B200NonlinearModeler(f, *args, 10.0 * np.eye(2), [-1.0], [1.0], None, None, 5, np.zeros(2), np.zeros(1))
    x0 = np.random.default_rng(0).uniform(-0.1, 0.1, (37, 2))
    res = mod.solve_batch(x0, np.zeros(2), np.zeros(1), want=("u",), method="linear")
    assert (res["status"] == mpc._lib.STATUS_DESIGN_FAILED).all() and (res["iters"] == 0).all()
    # the SQP solve of the same controller (terminal weight given) is unaffected
    res = mod.solve_batch(x0, np.zeros(2), np.zeros(1), want=("u",))
    assert set(np.unique(res["status"])) <= {1, -2}
    # a stabilisable sibling in the same process: f = [0.9 x1 + 0.3 u, 0.5 x2 + u]
    g = mpc.Fnn(np.eye(3), [(np.eye(3), np.zeros(3))], np.array([[0.9, 0.0, 0.3], [0.0, 0.5, 1.0]]), activation="identity")
    mod2 = mpc.B200NonlinearModeler(g, *args, None, [-1.0], [1.0], None, None, 5, np.zeros(2), np.zeros(1))
    r2 = mod2.solve_batch(x0, np.zeros(2), np.zeros(1), want=("u", "objective"), method="linear")
    assert (r2["status"] == 1).all()
    A = np.array([[0.9, 0.0], [0.0, 0.5]]); B = np.array([[0.3], [1.0]])
    c = mo.condense(A, B, Q, R, S, mo.dare(A, B, Q, R), 5, [-1.0], [1.0])
    for i in (0, 11):
        v, _ = mo.qp_exact(c, mo.pack_params(x0[i], np.zeros(2), np.zeros(1))[0])
        assert np.abs(r2["u"][i].ravel() - v).max() < 1e-7


def test_c_abi_error_behaviour_of_the_design_entries(mpc, qt):
    """Return codes + mpcb_last_error for the nonlinear / device-design entry points (no exceptions across the boundary)."""
    import ctypes as C
    L = mpc._lib.lib(); lib = mpc._lib
    m = load_nn_fixture("qt_fnn_tanh_model.json")
    mod = mpc.B200NonlinearModeler(to_chain(mpc, m), qt["Q"], qt["R"], qt["S"], None, qt["umin"], qt["umax"], None, None, 5, qt["x_ref"], qt["u_ref"])
    io = lib.BatchIO(); io.batch = 3
    for fn in (L.mpcb_solve_relinearized_batch, L.mpcb_solve_nmpc_batch):
        assert fn(mod._h, C.byref(io)) == -1 and b"x0" in L.mpcb_last_error()                    # missing inputs
        assert fn(None, C.byref(io)) == -1
    x0 = np.zeros((3, 4)); io.x0 = x0.ctypes.data; io.xref = x0.ctypes.data; io.uref = x0.ctypes.data; io.batch = 0
    assert L.mpcb_solve_relinearized_batch(mod._h, C.byref(io)) == -1 and b"batch" in L.mpcb_last_error()
    cio = lib.ClosedLoopIO(); cio.batch = 3; cio.steps = 0; cio.x0 = x0.ctypes.data; cio.xref = x0.ctypes.data; cio.uref = x0.ctypes.data
    assert L.mpcb_closed_loop_nmpc_batch(mod._h, C.byref(cio)) == -1 and b"steps" in L.mpcb_last_error()
    # batched Riccati equations
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    A = np.zeros((2, 4)); B = np.zeros((2, 2)); Q = np.eye(2); R = np.eye(1); P = np.zeros((2, 4))
    assert L.mpcb_dare_batch(0, 0, 2, 1, p(A), p(B), p(Q), p(R), p(P), None) == -1 and b"bad arguments" in L.mpcb_last_error()
    assert L.mpcb_dare_batch(99, 2, 2, 1, p(A), p(B), p(Q), p(R), p(P), None) == -1 and b"device" in L.mpcb_last_error()
    B7 = np.zeros((2, 14)); R7 = np.eye(7)
    assert L.mpcb_dare_batch(0, 2, 2, 7, p(A), p(B7), p(Q), p(R7), p(P), None) == -1 and b"nu <= 3 nx" in L.mpcb_last_error()
    # design-time refusals of the nonlinear controller
    with pytest.raises(mpc.MpcbError, match="not supported"):
        mpc.B200NonlinearModeler(to_chain(mpc, m), qt["Q"], qt["R"], qt["S"], None, qt["umin"], qt["umax"], None, None, 5, qt["x_ref"], qt["u_ref"], terminal="neighborhood")
    with pytest.raises(mpc.MpcbError, match="xmin"):
        mpc.B200NonlinearModeler(to_chain(mpc, m), qt["Q"], qt["R"], qt["S"], None, qt["umin"], qt["umax"], None, None, 5, qt["x_ref"], qt["u_ref"], state_constraint=True)
    with pytest.raises(mpc.MpcbError, match="128"):
        mpc.B200NonlinearModeler(to_chain(mpc, m), qt["Q"], qt["R"], qt["S"], None, qt["umin"], qt["umax"], None, None, 70, qt["x_ref"], qt["u_ref"])
