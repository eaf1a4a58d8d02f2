"""Parity at the sizes BASELINE.json names (configs[2], [3], [4]) -- the code paths whose defects only show at full size: the
streamed kernel's ticket scheduler and compaction at 8 192 x 896, the stage-wise kernel's tile staging at 16 384 x 400, the SQP
kernel's work queue at 4 096 problems.  EVERY problem is compared with the oracle twin (or, where the twin at full size would
take minutes on the CPU, a random quarter of the batch with the twin and all of it through solver-independent properties), a
random >= 1 % sample with the exact optimum / an independent solver."""
import numpy as np
import pytest

from conftest import assert_matches_twin, exact_sample, load_nn_fixture, qt_batch
from oracle import mpc_oracle as mo
from oracle import nn_oracle as no
from test_gpu_linear import OBJ_TOL, RES_TOL, U0_TOL, make_controller, oracle_condensed

pytestmark = pytest.mark.gpu


def lti64(seed=1, nx=64, nu=16):
    """random stable LTI, rng(1)."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((nx, nx)); A = 0.95 * G / np.abs(np.linalg.eigvals(G)).max(); B = rng.standard_normal((nx, nu)) / 8
    return A, B


def test_config3_lti64_terminal_equality_at_size(mpc):
    """configs[2]: nx = 64, nu = 16, H = 50, box-constrained with terminal LQR cost + terminal equality, batch 8 192, at the PARITY
    tolerance eps = 1e-7.  nz = 800, nt = 864 -> streamed DMMA GEMM kernel (the stage-wise form does not apply: general rows, nx = 64)."""
    from oracle import osqp_ref as orf
    A, B = lti64()
    nx, nu, H, n, eps, check = 64, 16, 50, 8192, 1e-7, 10
    umin, umax = -np.ones(nu), np.ones(nu)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-1e3 * np.ones(nx), 1e3 * np.ones(nx)), mpc.Hyperrectangle(umin, umax))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, mpc_solver="b200", mpc_terminal_ingredient="equality",
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=0.0, mpc_b200_max_iter=4000)
    m = C.tuning.modeler
    assert m.info.kernel == 2 and m.info.nz == 800 and m.info.mg == 64 and m.info.nt_pad == 896
    rng = np.random.default_rng(3)
    x0 = 1.0 * rng.standard_normal((n, nx)); xref = np.zeros(nx); uref = np.zeros(nu)      # s = 1: every problem feasible on this system (reported below)
    x0[:32] *= 1000.0; x0[32:64] *= 200.0      # ... plus 64 far-away states: the terminal equality is infeasible under the input box for all / most of them
    res = m.solve_batch(x0, xref, uref, want=("u", "u0", "objective", "x", "e_x"))
    st = res["status"]
    n_inf, n_cap = int((st == -3).sum()), int((st == -2).sum())
    print(f"config 3 at size: solved {int((st == 1).sum())}, infeasible {n_inf}, iteration cap {n_cap}, mean iters {res['iters'][st == 1].mean():.1f}, max {res['iters'].max()}")
    assert set(np.unique(st)) <= {1, -2, -3} and (st[64:] == 1).mean() >= 0.95 and n_inf >= 32
    ok = st == 1
    # solver-independent properties on EVERY solved problem: input box, terminal equality, residuals at the parity tolerance, recursion
    assert (res["u"][ok] >= umin - 1e-5).all() and (res["u"][ok] <= umax + 1e-5).all()
    assert np.abs(res["e_x"][ok][:, -1, :]).max() < 1e-5
    P = C.tuning.terminal_ingredient.P
    c = mo.condense(A, B, 100 * np.eye(nx), 0.1 * np.eye(nu), np.zeros((nu, nu)), P, H, umin, umax, terminal="equality")
    p = mo.pack_params(x0, xref, uref)
    # residuals: the primal one is in input units; the dual one is a gradient of a cost with |q|_inf ~ 1e3..1e4 here (Q = 100, lambda_max(Pc) =
    # 4.6e5), so it is measured against |q|_inf like OSQP's own relative criterion (for the quadruple tank |q| is O(10) and the absolute bound holds)
    qn = np.abs(p @ c.Lq.T).max(1)
    assert res["prim_res"][ok].max() < RES_TOL and (res["dual_res"][ok] / np.maximum(1.0, qn[ok])).max() < RES_TOL
    rec = mo.recover(c, res["u"].reshape(n, -1), p)
    assert np.abs(res["x"] - rec["x"]).max() < 1e-9 * max(1.0, np.abs(rec["x"]).max())
    assert np.abs(res["objective"] - rec["objective"]).max() <= 1e-10 * np.abs(rec["objective"]).max()
    # twin on a random quarter of the batch (+ the far-away block): statuses, iteration counts, solutions
    # (the far-away block separately: its stragglers run to the 4 000-iteration cap, which the numpy twin would make the whole sample pay)
    for sel, status_frac in ((np.arange(64), 0.9), (64 + np.sort(np.random.default_rng(0).choice(n - 64, 2048, replace=False)), 1.0)):
        tw = mo.admm_condensed(c, p[sel], mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=check, sigma=0.0, max_iter=4000))
        sub = {k: res[k][sel] for k in ("u", "status", "iters")}
        assert_matches_twin(sub, tw, tight=1e-7, loose=1e-4, min_same=0.9 if status_frac == 1.0 else 0.5, status_frac=status_frac, check=check)
    # exact optimum on a sample of the solved problems: u0 metric and objective
    idx = np.random.default_rng(1).choice(np.flatnonzero(ok), 12, replace=False)
    for i in idx:
        v, info = mo.qp_exact(c, p[i], v_init=res["u"][i].ravel())
        assert mo.u0_metric(res["u0"][i], v[:nu], umin, umax) < U0_TOL
        Jex = mo.recover(c, v[None], p[i][None])["objective"][0]
        assert abs(res["objective"][i] - Jex) <= OBJ_TOL * abs(Jex)
    # and the OSQP port on the reference's own sparse encoding (n = 12 192 variables) for two of them
    for i in idx[:2]:
        qp = mo.build_reference_qp(A, B, 100 * np.eye(nx), 0.1 * np.eye(nu), np.zeros((nu, nu)), P, H, xref, uref, x0[i], umin, umax, terminal="equality")
        w = orf.Workspace(orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u), orf.default_settings(eps_abs=1e-7, eps_rel=1e-7, max_iter=100000))
        r = w.solve(cold_start=True)
        assert r["status"] == 1
        u_sparse = r["x"][qp.idx["u"].T.ravel()]
        assert mo.u0_metric(res["u0"][i], u_sparse[:nu], umin, umax) < U0_TOL
        assert abs(r["obj"] - res["objective"][i]) <= 1e-5 * max(1.0, abs(r["obj"]))


def test_config3_lti64_tuned_step_size(mpc):
    """mpcb_tune_rho (kw `mpc_b200_rho_tune`): the step size picked on a 512-problem sample of the workload -- the batch-wide counterpart of
    OSQP's per-problem adaptive rho.  The automatic value sqrt(lambda_min lambda_max) suits batches with many active bounds; for this
    workload (few active bounds) a smaller one needs a third of the iterations.  Same optima: u0 and objective against the automatic-rho
    solve and the twin run at the tuned value."""
    A, B = lti64()
    nx, nu, H, n, eps, check = 64, 16, 50, 8192, 1e-7, 10
    umin, umax = -np.ones(nu), np.ones(nu)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-1e3 * np.ones(nx), 1e3 * np.ones(nx)), mpc.Hyperrectangle(umin, umax))
    x0 = np.random.default_rng(3).standard_normal((n, nx)); xref = np.zeros(nx); uref = np.zeros(nu)
    kw = dict(mpc_solver="b200", mpc_terminal_ingredient="equality", mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=0.0)
    Ca = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, **kw)
    Ct = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, mpc_b200_rho_tune=(x0[:512], xref, uref), **kw)
    ma, mt = Ca.tuning.modeler, Ct.tuning.modeler
    tun = mt.rho_tuning
    print("rho tuning:", [(round(r, 1), round(i, 1)) for r, i in zip(tun["candidates"], tun["mean_iters"])], "-> rho", round(tun["rho"], 1), "automatic", round(ma.info.rho, 1))
    assert abs(mt.info.rho - tun["rho"]) < 1e-9 * tun["rho"] and min(tun["mean_iters"]) == tun["mean_iters"][tun["candidates"].index(tun["rho"])]      # ("mean_iters": the group-max cost)
    ra = ma.solve_batch(x0, xref, uref, want=("u", "u0", "objective")); rt = mt.solve_batch(x0, xref, uref, want=("u", "u0", "objective"))
    assert (ra["status"] == 1).all() and (rt["status"] == 1).all()
    assert rt["iters"].mean() < 0.6 * ra["iters"].mean()
    assert mo.u0_metric(rt["u0"], ra["u0"], umin, umax).max() < U0_TOL
    assert (np.abs(rt["objective"] - ra["objective"]) <= OBJ_TOL * np.abs(ra["objective"])).all()
    c = mo.condense(A, B, 100 * np.eye(nx), 0.1 * np.eye(nu), np.zeros((nu, nu)), Ct.tuning.terminal_ingredient.P, H, umin, umax, terminal="equality")
    sel = np.sort(np.random.default_rng(0).choice(n, 1024, replace=False))
    tw = mo.admm_condensed(c, mo.pack_params(x0[sel], xref, uref), mo.AdmmSettings(rho=mt.info.rho, eps_abs=eps, eps_rel=eps, check_every=check, sigma=0.0))
    assert_matches_twin({k: rt[k][sel] for k in ("u", "status", "iters")}, tw, tight=1e-7, loose=1e-4, min_same=0.9, check=check)


@pytest.mark.parametrize("H", [100, 150, 200])
def test_config4_long_horizons_at_size(mpc, qt, H):
    """configs[3], long end of the horizon sweep at its batch of 16 384: the stage-wise kernel (automatic choice) against the twin
    on a random quarter of the batch (all of it at H = 100), against the streamed DMMA GEMM kernel on ALL problems, and against the
    exact optimum on a >= 1 % sample."""
    n, eps, check = 16384, 1e-7, 5
    x0, xref, uref = qt_batch(qt, n)
    out = {}
    for kern in (0, 2):
        C = make_controller(mpc, qt, H, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=0.0, mpc_b200_kernel=kern)
        m = C.tuning.modeler
        assert m.info.kernel == (4 if kern == 0 else 2)
        out[kern] = m.solve_batch(x0, xref, uref, want=("u", "u0", "objective"))
        rho = m.info.rho; P = C.tuning.terminal_ingredient.P
    a, b = out[0], out[2]
    assert (a["status"] == 1).all() and (b["status"] == 1).all()
    same = a["iters"] == b["iters"]
    assert same.mean() > 0.99 and np.abs(a["u"][same] - b["u"][same]).max() < 1e-9 and np.abs(a["u"] - b["u"]).max() < 1e-6
    assert np.abs(a["objective"] - b["objective"]).max() <= 1e-9 * np.abs(b["objective"]).max()
    assert a["prim_res"].max() < RES_TOL and a["dual_res"].max() < RES_TOL
    c = oracle_condensed(qt, H, P)
    p = mo.pack_params(x0, xref, uref)
    sel = np.arange(n) if H == 100 else np.sort(np.random.default_rng(H).choice(n, n // 4, replace=False))
    tw = mo.admm_condensed(c, p[sel], mo.AdmmSettings(rho=rho, eps_abs=eps, eps_rel=eps, check_every=check, sigma=0.0))
    for r in (a, b):
        assert_matches_twin({k: r[k][sel] for k in ("u", "status", "iters")}, tw, tight=1e-9, loose=1e-6, min_same=0.99, check=check)
    idx = exact_sample(n, frac=0.01, at_least=164, seed=H)
    ex = np.array([mo.qp_exact(c, p[i], v_init=a["u"][i].ravel())[0] for i in idx])
    assert mo.u0_metric(a["u0"][idx], ex[:, :2], qt["umin"], qt["umax"]).max() < U0_TOL
    Jex = mo.recover(c, ex, p[idx])["objective"]
    assert (np.abs(a["objective"][idx] - Jex) / np.maximum(np.abs(Jex), 1e-9)).max() < OBJ_TOL


@pytest.mark.parametrize("fixture", ["qt_resnet_model.json", "qt_fnn_tanh_model.json", "qt_resnet_swish_model.json", "qt_polynet_tanh_model.json",
                                     "qt_densenet_tanh_model.json", "qt_fnn_model.json"])
def test_config5_nmpc_at_size(mpc, qt, fixture):
    """configs[4]: NMPC, H = 20, batch 4 096 on every network fixture (the ResNet surrogate is the named one): every problem against
    the SQP twin, a >= 1 % sample against an independent L-BFGS-B solve with a KKT certificate."""
    from test_gpu_nmpc import make_system, scenario
    m = load_nn_fixture(fixture)
    H, n = 20, 4096
    C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                               mpc_programming_type="non_linear")
    mod = C.tuning.modeler
    x0, xref, uref = scenario(qt, n)
    res = mod.solve_batch(x0, xref, uref)
    d = mod.design()
    tw = no.nmpc_sqp(m, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0, xref, np.tile(uref, (n, 1)), d["rho"])
    relu = m.activation == "relu" and m.arch == "fnn"
    st_eq = res["status"] == tw["status"]
    assert st_eq.mean() > (0.97 if not relu else 0.93), st_eq.mean()
    ok = st_eq & (res["status"] == 1)
    print(f"{fixture}: status 1 {float((res['status'] == 1).mean()):.4f}, 2 {float((res['status'] == 2).mean()):.4f}, -2 {float((res['status'] == -2).mean()):.4f}")
    assert ok.mean() > (0.9 if not relu else 0.8)
    assert ((res["iters"] == tw["iters"]) | ~ok).mean() > 0.9
    assert np.abs(res["u"][ok] - tw["u"][ok]).max() < (5e-6 if not relu else 1e-4)
    assert np.abs(res["objective"][ok] - tw["objective"][ok]).max() <= (1e-9 if not relu else 1e-7) * np.abs(tw["objective"]).max()
    # every problem: the returned trajectory IS the network's rollout of the returned inputs, inputs inside the box, cost not above the start
    assert np.abs(res["x"] - no.rollout(m, x0, res["u"])).max() < 1e-12
    s1 = res["status"] == 1      # (a problem flagged 2 / -2 returns the x~ of a QP that was cut short: within the QP's current residual of the box only)
    assert (res["u"][s1] >= qt["umin"] - 2e-9).all() and (res["u"][s1] <= qt["umax"] + 2e-9).all()
    Hc = no.constant_hessian(2, H, qt["R"], qt["S"])
    J_init, _ = no.objective(m, qt["Q"], d["P"], Hc, np.clip(np.tile(uref, (n, H, 1)), qt["umin"], qt["umax"]), x0, xref, np.tile(uref, (n, 1)))
    assert (res["objective"] <= J_init + 1e-9 * np.abs(J_init)).all()
    if not relu:
        for i in np.random.default_rng(5).choice(np.flatnonzero(ok), 41, replace=False):
            u, J, k = no.nmpc_local_opt(m, qt["Q"], qt["R"], qt["S"], d["P"], H, qt["umin"], qt["umax"], x0[i], xref[i], uref, u_init=res["u"][i])
            assert k < 1e-5
            assert mo.u0_metric(res["u0"][i], u[0], qt["umin"], qt["umax"]) < U0_TOL
            assert abs(J - res["objective"][i]) <= OBJ_TOL * abs(J)
