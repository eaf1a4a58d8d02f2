"""GPU parity tests of the stage-wise (Riccati) kernel, mpc_b200_kernel = 4 (csrc/admm_riccati.cu): the same ADMM as the condensed
kernels with the x-update done by a cached Riccati sweep over the horizon (the reference's own stage-wise formulation,
linear.jl:48-60).  Checked against the condensed oracle twin on EVERY problem, against the exact optimum on a sample, and
against the condensed CUDA kernels."""
import numpy as np
import pytest

from conftest import assert_matches_twin, qt_batch
from oracle import mpc_oracle as mo
from test_gpu_linear import OBJ_TOL, RES_TOL, U0_TOL, make_controller, oracle_condensed

pytestmark = pytest.mark.gpu


def check_against_twin(res, tw, n, tight=1e-9, loose=1e-6):
    return assert_matches_twin(res, tw, tight=tight, loose=loose, min_same=0.99)[0]


@pytest.mark.parametrize("H,sigma,n", [(20, 0.0, 1500), (20, 1e-6, 700), (50, 0.0, 1500), (75, 0.0, 333), (100, 1e-6, 300), (7, 0.0, 65), (33, 0.0, 1)])
def test_riccati_matches_twin_and_exact(mpc, qt, H, sigma, n):
    eps, check = 1e-7, 5
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=sigma, mpc_b200_kernel=4)
    m = C.tuning.modeler
    assert m.info.kernel == 4
    x0, xref, uref = qt_batch(qt, n, seed=31)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P)
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=check, sigma=sigma))
    assert (res["status"] == 1).all()
    v = check_against_twin(res, tw, n)
    assert res["prim_res"].max() < RES_TOL and res["dual_res"].max() < RES_TOL
    assert np.abs(res["prim_res"] - tw["prim_res"]).max() < 1e-9
    rec = mo.recover(c, v, p)
    for k in ("x", "e_x", "u", "e_u"):
        assert np.abs(res[k] - rec[k]).max() < 1e-10, k
    ne = min(n, 48)
    ex = np.array([mo.qp_exact(c, p[i], v_init=tw["v"][i])[0] for i in range(ne)])
    assert mo.u0_metric(res["u0"][:ne], ex[:, :2], qt["umin"], qt["umax"]).max() < U0_TOL
    Jex = mo.recover(c, ex, p[:ne])["objective"]
    assert (np.abs(res["objective"][:ne] - Jex) / np.maximum(np.abs(Jex), 1e-9)).max() < OBJ_TOL


def test_riccati_equals_condensed_kernels(mpc, qt):
    """Three independent CUDA kernels, one algorithm: stage-wise vs on-chip (H = 20), vs shared-memory (H = 50), vs streamed (H = 70),
    incl. the duals."""
    n = 1200
    x0, xref, uref = qt_batch(qt, n, seed=32)
    for H, other in ((20, 1), (50, 3), (70, 2)):
        out = []
        for kern in (4, other):
            C = make_controller(mpc, qt, H, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_kernel=kern)
            out.append(C.tuning.modeler.solve_batch(x0, xref, uref, want=("u", "u0", "objective", "y")))
        a, b = out
        same = a["iters"] == b["iters"]
        assert same.mean() > 0.99 and np.array_equal(a["status"], b["status"])
        assert np.abs(a["u"][same] - b["u"][same]).max() < 1e-9 and np.abs(a["u"] - b["u"]).max() < 1e-6
        assert np.abs(a["y"][same] - b["y"][same]).max() < 1e-7
        assert np.abs(a["objective"] - b["objective"]).max() <= 1e-9 * np.abs(b["objective"]).max()


def test_riccati_warm_start_and_closed_loop(mpc, qt):
    H, n = 40, 500
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_kernel=4)
    m = C.tuning.modeler
    x0, xref, uref = qt_batch(qt, n, seed=33)
    r = m.solve_batch(x0, xref, uref, want=("u", "y"))
    w = m.solve_batch(x0, xref, uref, want=("u",), warm=(r["u"], r["y"]))
    assert (w["status"] == 1).all() and w["iters"].max() <= 10 and np.abs(w["u"] - r["u"]).max() < 2e-5
    # GPU-resident closed loop == host-driven loop (same kernel both ways)
    steps = 6
    dev = m.closed_loop(x0, xref, uref, steps, warm_start=True)
    x = x0.copy(); m.warm = None; xs = [x.copy()]
    for t in range(steps):
        mpc.update_initialization(C, x, references=(xref, uref))
        rr = mpc.calculate(C, warm_start=True, want=("u", "u0"))
        assert (rr["status"] == 1).all()
        x = xref + (x - xref) @ qt["A"].T + (rr["u0"] - uref) @ qt["B"].T
        xs.append(x.copy())
    assert np.abs(dev["x_traj"] - np.stack(xs, 1)).max() < 1e-9 and (dev["unsolved_steps"] == 0).all()


@pytest.mark.parametrize("nx,nu,H,sigma", [(2, 1, 45, 0.0), (3, 1, 30, 1e-6), (3, 2, 64, 0.0), (5, 3, 33, 1e-6), (6, 2, 40, 0.0), (6, 3, 21, 0.0), (8, 4, 40, 0.0)])
def test_riccati_random_systems(mpc, nx, nu, H, sigma):
    rng = np.random.default_rng(1000 * nx + 10 * nu + H)
    G = rng.standard_normal((nx, nx)); A = 0.9 * G / np.abs(np.linalg.eigvals(G)).max(); B = rng.standard_normal((nx, nu)) / 2
    umin, umax = -np.ones(nu), np.ones(nu)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-50 * np.ones(nx), 50 * np.ones(nx)), mpc.Hyperrectangle(umin, umax))
    n, eps = 333, 1e-7
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, mpc_solver="b200", mpc_Q=10.0, mpc_R=1.0,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_sigma=sigma, mpc_b200_max_iter=20000, mpc_b200_kernel=4)
    m = C.tuning.modeler
    x0 = 2.0 * rng.standard_normal((n, nx)); xref = 0.2 * rng.standard_normal((n, nx)); uref = 0.1 * rng.standard_normal((n, nu))
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = mo.condense(A, B, 10 * np.eye(nx), np.eye(nu), np.zeros((nu, nu)), C.tuning.terminal_ingredient.P, H, umin, umax)
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, sigma=sigma, max_iter=20000))
    v = check_against_twin(res, tw, n)
    rec = mo.recover(c, v, p)
    assert np.abs(res["x"] - rec["x"]).max() < 1e-9 * max(1.0, np.abs(rec["x"]).max())


def test_riccati_refuses_what_it_cannot_do(mpc, qt):
    with pytest.raises(mpc.MpcbError, match="Riccati"):
        make_controller(mpc, qt, 20, terminal="equality", mpc_b200_kernel=4)          # general rows
    with pytest.raises(mpc.MpcbError, match="Riccati"):
        make_controller(mpc, qt, 20, mpc_S=1.0, mpc_b200_kernel=4)                    # S term couples the stages' inputs


def test_tune_rho_small_sample_and_explicit_rho(mpc, qt):
    """mpcb_tune_rho on a sample too small for the timed path (scored by the group-maximum iteration count instead), with an explicit starting
    value: the candidates are centred on it, the winner is one of them, and the controller built with it solves the batch to the same optima
    as the automatic step size."""
    H, n = 100, 2000
    x0, xref, uref = qt_batch(qt, n, seed=44)
    kw = dict(mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=0.0)
    Ca = make_controller(mpc, qt, H, **kw)
    Ct = make_controller(mpc, qt, H, mpc_b200_rho=2.0, mpc_b200_rho_tune=(x0[:96], xref[:96], uref, 5), **kw)
    t = Ct.tuning.modeler.rho_tuning
    assert len(t["candidates"]) == 5 and abs(t["candidates"][2] - 2.0) < 1e-12 and t["rho"] in t["candidates"]
    assert all(s > 1.0 for s in t["mean_iters"])                      # iteration-count scores (a timed score would be a fraction of a millisecond)
    assert abs(Ct.tuning.modeler.info.rho - t["rho"]) < 1e-12
    ra = Ca.tuning.modeler.solve_batch(x0, xref, uref, want=("u0", "objective")); rt = Ct.tuning.modeler.solve_batch(x0, xref, uref, want=("u0", "objective"))
    assert (ra["status"] == 1).all() and (rt["status"] == 1).all()
    assert mo.u0_metric(rt["u0"], ra["u0"], qt["umin"], qt["umax"]).max() < U0_TOL
    assert (np.abs(rt["objective"] - ra["objective"]) <= OBJ_TOL * np.abs(ra["objective"])).all()
