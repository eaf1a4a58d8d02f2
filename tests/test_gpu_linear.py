"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

from conftest import assert_matches_twin, exact_sample, qt_batch
from oracle import mpc_oracle as mo

pytestmark = pytest.mark.gpu

U0_TOL = 1e-4        # north_star: optimal u0 within 1e-4 (metric of oracle.u0_metric)
OBJ_TOL = 1e-6       # objective within 1e-6 relative
RES_TOL = 1e-5       # primal/dual residual <= 1e-5


def make_controller(mpc, qt, H, terminal="none", **kw):
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(qt["xmin"], qt["xmax"]),
                                                      mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    return mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                                  mpc_terminal_ingredient=terminal, **kw)


def oracle_condensed(qt, H, P, terminal="none", state_constraint=False, S=None):
    return mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"] if S is None else S, P, H, qt["umin"], qt["umax"], qt["xmin"], qt["xmax"],
                       state_constraint=state_constraint, terminal=terminal)


def test_design_matches_oracle(mpc, qt):
    C = make_controller(mpc, qt, 20)
    d = C.tuning.modeler.design()
    c = oracle_condensed(qt, 20, C.tuning.terminal_ingredient.P)
    assert np.allclose(d["Pc"], c.Pc, rtol=1e-12, atol=1e-10)
    assert np.allclose(d["Lq"], c.Lq, rtol=1e-12, atol=1e-10)
    T, _, _, rho = mo.admm_matrices(c, mo.AdmmSettings(rho=C.tuning.modeler.info.rho))
    assert np.allclose(d["T"], T, rtol=1e-10, atol=1e-12)
    assert abs(C.tuning.modeler.info.rho - mo.auto_rho(c.Pc)) < 1e-3 * rho      # heuristic value: power / inverse iteration accuracy is enough


def test_kat_lqr_single_problem(mpc, qt):
    """Known answer: no bound is active from x0 = 0.6, so u0* is the LQR law for every horizon."""
    for H in (5, 20):
        C = make_controller(mpc, qt, H, mpc_b200_eps_abs=1e-8, mpc_b200_eps_rel=1e-8, mpc_b200_check_every=5)
        mpc.update_initialization(C, qt["x0"])
        res = mpc.calculate(C)
        assert res["status"][0] == 1
        assert np.allclose(C.computation_results.u[:, 0], [2.75594127, 2.95507466], atol=2e-6)
        assert C.computation_results.x.shape == (4, H + 1) and C.computation_results.e_u.shape == (2, H)
        assert np.allclose(C.computation_results.x[:, 0], qt["x0"])
        assert np.allclose(C.computation_results.e_x, C.computation_results.x - 0.65)
        assert np.allclose(C.computation_results.e_u, C.computation_results.u - 1.2)


@pytest.mark.parametrize("H,eps,check,sigma", [(20, 1e-7, 5, 0.0), (20, 1e-7, 5, 1e-6), (20, 1e-6, 5, 1e-6), (20, 1e-3, 25, 1e-6), (5, 1e-7, 5, 0.0),
                                                (10, 1e-5, 4, 1e-6), (13, 1e-7, 5, 0.0), (7, 1e-7, 3, 1e-6)])
def test_batch_matches_twin_and_exact(mpc, qt, H, eps, check, sigma):
    n = 2048
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=sigma)
    m = C.tuning.modeler
    x0, xref, uref = qt_batch(qt, n)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P)
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=check, sigma=sigma))
    # same algorithm, same iteration counts (summation order differs -> allow a vanishing fraction of off-by-one-check); EVERY problem
    # is compared: to round-off where the counts agree, to the solver accuracy where a borderline check flipped
    assert (res["status"] == 1).all()
    v, same = assert_matches_twin(res, tw, tight=1e-9, loose=max(1e-6, 30 * eps), min_same=0.995, check=check)
    assert np.abs(res["prim_res"][same] - tw["prim_res"][same]).max() < 1e-9
    rec = mo.recover(c, v, p)
    for k in ("x", "e_x", "u", "e_u"):
        assert np.abs(res[k] - rec[k]).max() < 1e-11, k
    assert np.abs(res["objective"] - rec["objective"]).max() <= 1e-11 * np.abs(rec["objective"]).max()
    assert np.array_equal(res["u0"], res["u"][:, 0, :])
    if eps <= 1e-6:   # against the exact optimum, on a random sample of the batch
        idx = exact_sample(n, frac=0.1, at_least=128, seed=H)
        ex = np.array([mo.qp_exact(c, p[i], v_init=tw["v"][i])[0] for i in idx])
        assert mo.u0_metric(res["u0"][idx], ex[:, :2], qt["umin"], qt["umax"]).max() < U0_TOL
        if eps <= 1e-7:   # the parity settings: eps_abs = eps_rel = 1e-7
            assert res["prim_res"].max() < RES_TOL and res["dual_res"].max() < RES_TOL
            Jex = mo.recover(c, ex, p[idx])["objective"]
            assert (np.abs(res["objective"][idx] - Jex) / np.maximum(np.abs(Jex), 1e-9)).max() < OBJ_TOL


def test_terminal_equality_matches_twin_and_exact(mpc, qt):
    H, n, eps = 10, 1024, 1e-7
    C = make_controller(mpc, qt, H, terminal="equality", mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=10,
                        mpc_b200_max_iter=20000)
    m = C.tuning.modeler
    assert m.info.mg == 4 and m.info.nt == 24
    rng = np.random.default_rng(3)
    xref = np.tile(qt["x_ref"], (n, 1))
    x0 = xref + 0.002 * rng.standard_normal((n, 4))     # small deviations: the terminal equality is mostly feasible
    mpc.update_initialization(C, x0, references=(xref, qt["u_ref"]))
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P, terminal="equality")
    p = mo.pack_params(x0, xref, qt["u_ref"])
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=10, max_iter=20000))
    v, same = assert_matches_twin(res, tw, tight=1e-8, loose=1e-5, min_same=0.99, status_frac=0.995, check=10)
    ok = res["status"] == 1
    assert ok.mean() > 0.5
    assert np.abs(res["e_x"][ok][:, -1, :]).max() < 1e-5       # terminal deviation driven to zero
    idx = np.random.default_rng(0).choice(np.flatnonzero(ok), 64, replace=False)
    ex = np.array([mo.qp_exact(c, p[i], v_init=tw["v"][i])[0] for i in idx])
    assert mo.u0_metric(res["u0"][idx], ex[:, :2], qt["umin"], qt["umax"]).max() < U0_TOL


def test_terminal_equality_infeasible_is_flagged(mpc, qt):
    """Far-away initial states cannot reach the reference in H steps within the input box: OSQP's primal
    infeasibility certificate must fire per problem (status -3), not crash."""
    H, n = 5, 256
    C = make_controller(mpc, qt, H, terminal="equality", mpc_b200_check_every=10, mpc_b200_max_iter=4000)
    rng = np.random.default_rng(4)
    x0 = rng.uniform(qt["xmin"], qt["xmax"], (n, 4))
    mpc.update_initialization(C, x0)
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P, terminal="equality")
    p = mo.pack_params(x0, qt["x_ref"], qt["u_ref"])
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=C.tuning.modeler.info.rho, check_every=10, max_iter=4000))
    assert (res["status"] == tw["status"]).mean() > 0.98
    assert (res["status"] == -3).sum() > 0


def test_warm_start_closed_loop(mpc, qt):
    """Config 1: closed loop from x0 = 0.6 with warm starts (OSQP's default behaviour on a persistent model)."""
    H = 20
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=1e-6, mpc_b200_eps_rel=1e-6, mpc_b200_check_every=5)
    x = qt["x0"].copy()
    iters = []
    for k in range(30):
        mpc.update_initialization(C, x)
        res = mpc.calculate(C, warm_start=True)
        assert res["status"][0] == 1
        iters.append(int(res["iters"][0]))
        u0 = C.computation_results.u[:, 0]
        assert (u0 >= qt["umin"] - 1e-6).all() and (u0 <= qt["umax"] + 1e-6).all()
        x = qt["x_ref"] + qt["A"] @ (x - qt["x_ref"]) + qt["B"] @ (u0 - qt["u_ref"])
    assert np.abs(x - qt["x_ref"]).max() < np.abs(qt["x0"] - qt["x_ref"]).max()
    assert np.mean(iters[1:]) <= iters[0]       # warm starts do not cost more than the cold first solve


def test_input_rate_weight_S(mpc, qt):
    H, n, eps = 10, 512, 1e-6
    C = make_controller(mpc, qt, H, mpc_S=5.0, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5)
    x0, xref, uref = qt_batch(qt, n, seed=5)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P, S=5.0 * np.eye(2))
    p = mo.pack_params(x0, xref, uref)
    ex = np.array([mo.qp_exact(c, p[i])[0] for i in range(64)])
    assert mo.u0_metric(res["u0"][:64], ex[:, :2], qt["umin"], qt["umax"]).max() < U0_TOL
    rec = mo.recover(c, res["u"].reshape(n, -1), p)
    assert np.abs(res["objective"] - rec["objective"]).max() <= 1e-10 * np.abs(rec["objective"]).max()


def test_empty_and_ragged_batches(mpc, qt):
    C = make_controller(mpc, qt, 20, mpc_b200_eps_abs=1e-6, mpc_b200_eps_rel=1e-6, mpc_b200_check_every=5)
    m = C.tuning.modeler
    with pytest.raises(mpc.MpcbError):
        m.solve_batch(np.zeros((0, 4)), qt["x_ref"], qt["u_ref"])
    ref = None
    X0, XREF, uref = qt_batch(qt, 3000, seed=11)
    for n in (1, 7, 8, 9, 33, 1000, 1184, 1185, 3000):      # not multiples of the 8 problem slots a warp holds; up to 1 184 (8 per SM) on the cooperative kernel, beyond on the slot kernel
        x0, xref = X0[:n], XREF[:n]
        r = m.solve_batch(x0, xref, uref)
        assert (r["status"] == 1).all() and r["u"].shape == (n, 20, 2)
        if ref is None: ref = r["u"][0].copy()
        assert np.array_equal(r["u"][0], ref)              # a problem's answer does not depend on its batch -- nor on the kernel that batch size selects
        if n == 1184: small = {k: r[k][:1000].copy() for k in ("u", "x", "e_x", "e_u", "objective", "iters")}
        if n == 3000:
            for k, v in small.items(): assert np.array_equal(r[k][:1000], v), k      # cooperative + direct-recover path == slot + tiled-recover path, bit for bit


# ------------------------------------------------------------------------------------------------------------------
# streamed kernel (nt > 64, or forced with mpc_b200_kernel = 2)
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,force,sigma,expect", [(20, 2, 0.0, 2), (20, 2, 1e-6, 2), (50, 2, 0.0, 2), (35, 2, 1e-6, 2), (70, 2, 0.0, 2),
                                                   # beyond the on-chip range the stage-wise (Riccati) kernel is the automatic choice for box-only problems
                                                   (70, 0, 0.0, 4), (60, 0, 1e-6, 4),
                                                   # shared-memory resident kernel (box-only, 64 < nz <= 120 and the state fits 227 KB):
                                                   # picked automatically; nz = 120 with sigma > 0 (4 state arrays) does not fit -> stage-wise / streamed
                                                   (50, 0, 0.0, 3), (35, 0, 1e-6, 3), (33, 3, 0.0, 3), (60, 0, 0.0, 3), (60, 2, 1e-6, 2), (47, 0, 1e-6, 3)])
def test_streamed_matches_twin_and_exact(mpc, qt, H, force, sigma, expect):
    n, eps, check = 700, 1e-7, 5          # 700: not a multiple of the 128-row GEMM tile
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=sigma,
                        mpc_b200_kernel=force)
    m = C.tuning.modeler
    assert m.info.kernel == expect
    x0, xref, uref = qt_batch(qt, n, seed=21)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P)
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=check, sigma=sigma))
    assert (res["status"] == 1).all()
    v, same = assert_matches_twin(res, tw, tight=1e-9, loose=1e-6, min_same=0.99, check=check)
    idx = exact_sample(n, at_least=64, seed=H)
    ex = np.array([mo.qp_exact(c, p[i], v_init=tw["v"][i])[0] for i in idx])
    assert mo.u0_metric(res["u0"][idx], ex[:, :2], qt["umin"], qt["umax"]).max() < U0_TOL
    rec = mo.recover(c, v, p)
    assert np.abs(res["x"] - rec["x"]).max() < 1e-10 and np.abs(res["objective"] - rec["objective"]).max() <= 1e-10 * np.abs(rec["objective"]).max()


def test_streamed_equals_onchip(mpc, qt):
    """Two independent kernels, one algorithm: identical iteration counts and (to rounding) identical solutions."""
    n = 5000          # above the small-batch thresholds (8 problems per SM, 32 for the shared-memory resident kernels): the slot kernels
    x0, xref, uref = qt_batch(qt, n, seed=22)
    out = []
    for kern in (1, 2):
        C = make_controller(mpc, qt, 20, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_kernel=kern)
        mpc.update_initialization(C, x0, references=(xref, uref))
        out.append(mpc.calculate(C))
    assert (out[0]["iters"] == out[1]["iters"]).mean() > 0.995
    same = out[0]["iters"] == out[1]["iters"]
    assert np.abs(out[0]["u"][same] - out[1]["u"][same]).max() < 1e-10
    # and the shared-memory kernel against the streamed one at a horizon both accept, incl. duals and a warm-started second solve
    res = []
    for kern in (3, 2):
        C = make_controller(mpc, qt, 40, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_kernel=kern)
        m = C.tuning.modeler
        r = m.solve_batch(x0, xref, uref, want=("u", "u0", "objective", "y"))
        res.append(r)
        if kern == 3:
            w = m.solve_batch(x0, xref, uref, want=("u",), warm=(r["u"], r["y"]))
            assert (w["status"] == 1).all() and w["iters"].max() <= 10 and np.abs(w["u"] - r["u"]).max() < 2e-5       # both within eps of the optimum
    same = res[0]["iters"] == res[1]["iters"]
    assert same.mean() > 0.995 and np.abs(res[0]["u"][same] - res[1]["u"][same]).max() < 1e-10
    assert np.abs(res[0]["y"][same] - res[1]["y"][same]).max() < 1e-8


@pytest.mark.parametrize("kernel", [2, 0])
def test_streamed_terminal_equality(mpc, qt, kernel):
    """Terminal-equality rows at nt = 84: the streamed kernel (forced) and, chosen automatically since round 2, the shared-memory resident
    general-row kernel (admm_smemg.cuh) -- both against the twin."""
    H, n, eps = 40, 300, 1e-7
    C = make_controller(mpc, qt, H, terminal="equality", mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=10, mpc_b200_max_iter=6000,
                        mpc_b200_rho=10.0, mpc_b200_kernel=kernel)
    m = C.tuning.modeler
    assert m.info.kernel == (2 if kernel == 2 else 3) and m.info.mg == 4 and m.info.nt == 84
    rng = np.random.default_rng(5)
    xref = np.tile(qt["x_ref"], (n, 1))
    x0 = xref + 0.05 * rng.standard_normal((n, 4))
    mpc.update_initialization(C, x0, references=(xref, qt["u_ref"]))
    res = mpc.calculate(C)
    c = oracle_condensed(qt, H, C.tuning.terminal_ingredient.P, terminal="equality")
    p = mo.pack_params(x0, xref, qt["u_ref"])
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=10, max_iter=6000))
    v, same = assert_matches_twin(res, tw, tight=1e-7, loose=1e-4, min_same=0.5, status_frac=0.98, check=10)
    ok = res["status"] == 1
    assert np.abs(res["e_x"][ok][:, -1, :]).max() < 1e-5
    assert set(np.unique(res["status"])) <= {1, -2, -3}


def test_random_lti_generic_recover(mpc):
    """nx=5, nu=3: no register-resident recover specialisation -> generic recover kernel; nz = 36 -> on-chip NT = 40 with padding rows."""
    rng = np.random.default_rng(9)
    nx, nu, H, n = 5, 3, 12, 400
    G = rng.standard_normal((nx, nx)); A = 0.9 * G / np.abs(np.linalg.eigvals(G)).max(); B = rng.standard_normal((nx, nu)) / 2
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-5 * np.ones(nx), 5 * np.ones(nx)), mpc.Hyperrectangle(-np.ones(nu), np.ones(nu)))
    eps = 1e-7
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, mpc_solver="b200", mpc_Q=10.0, mpc_R=1.0,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5)
    m = C.tuning.modeler
    assert m.info.nz == 36 and m.info.nt_pad == 40
    x0 = 2.0 * rng.standard_normal((n, nx)); xref = 0.2 * rng.standard_normal((n, nx)); uref = 0.1 * rng.standard_normal((n, nu))
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = mo.condense(A, B, 10 * np.eye(nx), np.eye(nu), np.zeros((nu, nu)), C.tuning.terminal_ingredient.P, H, -np.ones(nu), np.ones(nu))
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5))
    assert (res["status"] == 1).all()
    v, same = assert_matches_twin(res, tw, tight=1e-9, loose=1e-6, min_same=0.99)
    rec = mo.recover(c, v, p)
    for k in ("x", "e_x", "u", "e_u"):
        assert np.abs(res[k] - rec[k]).max() < 1e-10, k
    assert np.abs(res["objective"] - rec["objective"]).max() <= 1e-10 * np.abs(rec["objective"]).max()
    ex = np.array([mo.qp_exact(c, p[i], v_init=tw["v"][i])[0] for i in range(48)])
    assert mo.u0_metric(res["u0"][:48], ex[:, :nu], -np.ones(nu), np.ones(nu)).max() < U0_TOL


def test_pipelined_host_entry_equals_plain(mpc, qt):
    """Large batches in page-locked memory take the chunked, download-overlapped path of mpcb_solve_linear_batch; the
    per-problem results must be bit-identical to the plain path (pageable arrays), and the one-problem zero-copy path
    must agree with both."""
    import ctypes
    n, H = 20000, 20
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=0.0)
    m = C.tuning.modeler
    x0, xref, uref = qt_batch(qt, n)
    plain = m.solve_batch(x0, xref, uref)
    assert m.timing()["chunks"] == 1
    L = mpc._lib.lib()
    shapes = {"u": (n, H, 2), "e_u": (n, H, 2), "x": (n, H + 1, 4), "e_x": (n, H + 1, 4), "u0": (n, 2), "objective": (n,), "prim_res": (n,), "dual_res": (n,)}
    ptrs = []

    def pinned(shape, dtype=np.float64):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = L.mpcb_alloc_pinned(nbytes); assert p
        ptrs.append(p)
        return np.frombuffer((ctypes.c_char * nbytes).from_address(p), dtype=dtype).reshape(shape)

    try:
        out = {k: pinned(s) for k, s in shapes.items()}
        out["status"] = pinned((n,), np.int32); out["iters"] = pinned((n,), np.int32)
        px0 = pinned(x0.shape); px0[:] = x0
        pxr = pinned(xref.shape); pxr[:] = xref
        pur = pinned(uref.shape); pur[:] = uref
        m.solve_batch(px0, pxr, pur, out=out)
        assert m.timing()["chunks"] > 1
        for k in list(shapes) + ["status", "iters"]:
            assert np.array_equal(out[k], plain[k]), k
        one = m.solve_batch(x0[:3], xref[:3], uref)          # zero-copy small path
        for k in ("u", "x", "objective", "iters"):
            assert np.array_equal(one[k], plain[k][:3]), k
    finally:
        for p in ptrs: L.mpcb_free_pinned(p)


@pytest.mark.parametrize("H", [10, 20])
def test_state_constraint_rows(mpc, qt, H):
    """kw `mpc_state_constraint` (linear.jl:62-70): state-box rows on x[:,2..H+1] become general inequality rows
    (H = 10: nt = 60 and H = 20: nt = 120, both on the shared-memory resident general-row kernel).  The box is tightened to
    [0.55, 0.75] and the references sit partly beyond it, so the optimal trajectories press against the state bounds.
    CUDA vs the condensed twin, and vs the OSQP port run on the reference's own sparse formulation (independent encoding
    and solver) at tight tolerance."""
    from oracle import osqp_ref as orf
    n, eps = 256, 1e-7
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_state_constraint=True,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000)
    m = C.tuning.modeler
    assert m.info.mg == 4 * H and m.info.kernel == 3      # nt = 60 and nt = 120: shared-memory resident general-row kernel (round 2; before: register-resident / streamed)
    rng = np.random.default_rng(5)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    mpc.update_initialization(C, x0, references=(xref, qt["u_ref"]))
    res = mpc.calculate(C)
    P = C.tuning.terminal_ingredient.P
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], xmin, xmax, state_constraint=True)
    p = mo.pack_params(x0, xref, qt["u_ref"])
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000))
    assert (res["status"] == 1).all() and (tw["status"] == 1).all()
    assert (res["iters"] == tw["iters"]).mean() > 0.97 and res["iters"].max() < 2000
    assert np.abs(res["u"].reshape(n, -1) - tw["v"]).max() < 1e-7
    xs = res["x"]
    touching = (xs[:, 1:] > xmax - 1e-6).any(axis=(1, 2))
    assert touching.sum() >= n // 4                                   # the rows are really exercised
    assert (xs[:, 1:] <= xmax + 1e-5).all() and (xs[:, 1:] >= xmin - 1e-5).all()
    for i in np.flatnonzero(touching)[:4]:                            # independent check on the reference's sparse model
        qp = mo.build_reference_qp(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, xref[i], qt["u_ref"], x0[i], qt["umin"], qt["umax"], xmin, xmax,
                                   state_constraint=True)
        w = orf.Workspace(orf.Problem(qp.P, qp.q, qp.A, qp.l, qp.u), orf.default_settings(eps_abs=1e-8, eps_rel=1e-8, max_iter=200000))
        r = w.solve(cold_start=True)
        assert r["status"] == 1
        u_sparse = r["x"][qp.idx["u"].T.ravel()]
        assert mo.u0_metric(res["u0"][i], u_sparse[:2], qt["umin"], qt["umax"]) < U0_TOL
        assert abs(r["obj"] - res["objective"][i]) <= 1e-5 * max(1.0, abs(r["obj"]))


def test_closed_loop_on_gpu_equals_host_loop(mpc, qt):
    """the update_initialization! -> calculate! -> apply u[:,1] loop kept on the device for a batch of plants
    must reproduce the same loop driven from the host (one batched calculate! per step, warm-started the same way)."""
    H, n, steps = 20, 300, 12
    C = make_controller(mpc, qt, H, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=0.0)
    m = C.tuning.modeler
    x0, xref, uref = qt_batch(qt, n)
    for warm in (True, False):
        dev = m.closed_loop(x0, xref, uref, steps, warm_start=warm)
        x = x0.copy(); m.warm = None
        xs, us, its = [x.copy()], [], np.zeros(n, np.int64)
        for t in range(steps):
            mpc.update_initialization(C, x, references=(xref, uref))
            r = mpc.calculate(C, warm_start=warm, want=("u", "u0"))
            assert (r["status"] == 1).all()
            its += r["iters"]
            x = xref + (x - xref) @ qt["A"].T + (r["u0"] - uref) @ qt["B"].T
            xs.append(x.copy()); us.append(r["u0"].copy())
        xs = np.stack(xs, 1); us = np.stack(us, 1)
        assert np.abs(dev["x_traj"] - xs).max() < 1e-9 and np.abs(dev["u_traj"] - us).max() < 1e-8
        assert (dev["unsolved_steps"] == 0).all() and np.abs(dev["iters_total"] - its).max() <= 5 * 2      # a borderline check may flip on 1e-16 differences
        if warm: it_warm = dev["iters_total"].mean()
        else: assert it_warm < dev["iters_total"].mean()                      # the (unshifted, OSQP-style) warm start pays
    # the regulated plants approach their references
    assert np.abs(dev["x_traj"][:, -1] - xref).max() < np.abs(x0 - xref).max()


@pytest.mark.parametrize("n", [333, 5000])      # below / above the small-batch thresholds (8 problems per SM, 32 for the shared-memory resident kernels): cooperative / slot kernels
@pytest.mark.parametrize("nx,nu,H,terminal,sigma,S_w", [
    (2, 1, 7, "none", 0.0, 0.0),        # nz = 7: odd, scalar stores, recover_small<2,1>
    (3, 1, 30, "none", 1e-6, 0.0),      # odd nx: 8-byte cooperative stores in recover
    (3, 2, 11, "equality", 1e-6, 0.0),  # on-chip kernel with general rows, nt = 25 -> padded 32
    (6, 3, 9, "none", 0.0, 2.0),        # nz = 27 odd, S term
    (6, 2, 40, "none", 0.0, 0.0),       # nz = 80 -> shared-memory kernel
    (5, 3, 33, "none", 1e-6, 0.0),      # nz = 99 odd -> shared-memory kernel, generic recover
    (8, 4, 40, "none", 0.0, 0.0),       # nz = 160 -> streamed
    (7, 3, 20, "equality", 0.0, 0.0),   # nt = 67 with general rows -> streamed, generic recover
    (10, 5, 12, "none", 0.0, 0.0),      # nz = 60, generic recover with wider states
    (20, 3, 8, "none", 0.0, 1.5),       # nx >= 16: the cooperative wide recover kernel (one thread group per problem), with the S term
    (33, 2, 10, "none", 1e-6, 0.0),     # nx = 33: thread groups of 64 with 31 idle lanes
])
def test_random_systems_sweep(mpc, nx, nu, H, terminal, sigma, S_w, n):
    """Odd sizes, every kernel path, every recover path: random stable systems, CUDA vs the condensed twin (iteration counts,
    solutions) and vs the result reconstruction of the oracle."""
    rng = np.random.default_rng(100 * nx + 10 * nu + H)
    G = rng.standard_normal((nx, nx)); A = 0.9 * G / np.abs(np.linalg.eigvals(G)).max(); B = rng.standard_normal((nx, nu)) / 2
    umin, umax = -np.ones(nu), np.ones(nu)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-50 * np.ones(nx), 50 * np.ones(nx)), mpc.Hyperrectangle(umin, umax))
    eps = 1e-7
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, mpc_solver="b200", mpc_Q=10.0, mpc_R=1.0, mpc_S=S_w,
                               mpc_terminal_ingredient=terminal, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_sigma=sigma,
                               mpc_b200_max_iter=20000)
    m = C.tuning.modeler
    scale = 0.05 if terminal == "equality" else 2.0
    x0 = scale * rng.standard_normal((n, nx)); xref = 0.1 * scale * rng.standard_normal((n, nx)); uref = 0.1 * rng.standard_normal((n, nu))
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = mo.condense(A, B, 10 * np.eye(nx), np.eye(nu), S_w * np.eye(nu), C.tuning.terminal_ingredient.P, H, umin, umax, terminal=terminal)
    p = mo.pack_params(x0, xref, uref)
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, sigma=sigma, max_iter=20000))
    v, same = assert_matches_twin(res, tw, tight=1e-9 if terminal == "none" else 1e-7, loose=1e-6 if terminal == "none" else 1e-4,
                                  min_same=0.9 if terminal == "none" else 0.5, status_frac=0.98)
    rec = mo.recover(c, v, p)
    for k in ("x", "e_x", "u", "e_u"):
        assert np.abs(res[k] - rec[k]).max() < 1e-9 * max(1.0, np.abs(rec[k]).max()), k
    assert np.abs(res["objective"] - rec["objective"]).max() <= 1e-10 * np.abs(rec["objective"]).max()
    assert np.array_equal(res["u0"], res["u"][:, 0])



def test_contractive_terminal_set(mpc, qt):
    """mpc_terminal_ingredient = "contractive" (design_mpc.jl:333-340): e_H' e_H <= 0.9 e_0' e_0, a quadratic constraint the
    reference needs SCIP / Ipopt for; here a ball projection inside the ADMM.  Short horizon and a lazy controller (Q = 1,
    R = 10) make the ball active.  CUDA vs twin, and vs an independent SLSQP solve of the QCQP."""
    from scipy.optimize import minimize
    H, n, eps = 3, 512, 1e-8
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(qt["xmin"], qt["xmax"]), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_Q=1.0, mpc_R=10.0,
                               mpc_terminal_ingredient="contractive", mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000)
    m = C.tuning.modeler
    assert m.info.mg == 4 and m.info.kernel == 1
    rng = np.random.default_rng(3)
    xref = rng.uniform(0.5, 0.9, (n, 4)); x0 = xref + 0.15 * rng.standard_normal((n, 4))
    mpc.update_initialization(C, x0, references=(xref, qt["u_ref"]))
    res = mpc.calculate(C)
    c = mo.condense(qt["A"], qt["B"], np.eye(4), 10 * np.eye(2), qt["S"], C.tuning.terminal_ingredient.P, H, qt["umin"], qt["umax"], terminal="contractive")
    p = mo.pack_params(x0, xref, qt["u_ref"])
    tw = mo.admm_condensed(c, p, mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000))
    assert (res["status"] == 1).all() and (tw["status"] == 1).all()
    assert (res["iters"] == tw["iters"]).mean() > 0.98 and np.abs(res["u"].reshape(n, -1) - tw["v"]).max() < 1e-8
    e0 = np.linalg.norm(res["e_x"][:, 0], axis=1); eH = np.linalg.norm(res["e_x"][:, H], axis=1)
    assert (eH <= np.sqrt(0.9) * e0 + 1e-7).all()
    active = np.abs(eH - np.sqrt(0.9) * e0) < 1e-6
    assert active.sum() >= n // 10                                       # the ball is really active on part of the batch
    for i in np.flatnonzero(active)[:4]:
        q = c.Lq @ p[i]; b = c.Lb @ p[i]; r2 = 0.9 * e0[i] ** 2
        r = minimize(lambda v: (0.5 * v @ c.Pc @ v + q @ v, c.Pc @ v + q), res["u"][i].ravel(), jac=True, method="SLSQP", bounds=list(zip(c.lb, c.ub)),
                     constraints=[{"type": "ineq", "fun": lambda v: r2 - np.sum((c.G @ v - b) ** 2), "jac": lambda v: -2 * (c.G @ v - b) @ c.G}],
                     options={"maxiter": 1000, "ftol": 1e-16})
        assert mo.u0_metric(res["u0"][i], r.x[:2], qt["umin"], qt["umax"]) < U0_TOL
        v = res["u"][i].ravel()
        assert abs(r.fun - (0.5 * v @ c.Pc @ v + q @ v)) <= OBJ_TOL * max(1.0, abs(r.fun))
    # too large for the on-chip kernels (nt > 120) -> refused, not approximated
    with pytest.raises(mpc.MpcbError):
        make_controller(mpc, qt, 60, terminal="contractive")


def test_contractive_terminal_set_shared_memory_kernel(mpc):
    """The ball projection on the shared-memory resident general-row kernel (H = 40: nt = 84), against the twin.  A slow plant (poles at
    0.998 .. 0.9995: A^40 alone does not contract by sqrt(0.9)) with an expensive input makes the ball active on most of the batch."""
    H, n, eps = 40, 256, 1e-8
    rng = np.random.default_rng(12)
    A = np.diag([0.999, 0.9985, 0.9995, 0.998]); A[0, 1] = 0.001; A[2, 3] = -0.0005
    B = 0.01 * rng.standard_normal((4, 2))
    umin, umax = -np.ones(2), np.ones(2)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-10 * np.ones(4), 10 * np.ones(4)), mpc.Hyperrectangle(umin, umax))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, [0.0] * 4, [0.0] * 2, mpc_solver="b200", mpc_Q=1.0, mpc_R=1000.0,
                               mpc_terminal_ingredient="contractive", mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000)
    m = C.tuning.modeler
    assert m.info.mg == 4 and m.info.nt == 84 and m.info.kernel == 3
    x0 = 0.5 * rng.standard_normal((n, 4)); xref = np.zeros((n, 4)); uref = np.zeros(2)
    mpc.update_initialization(C, x0, references=(xref, uref))
    res = mpc.calculate(C)
    c = mo.condense(A, B, np.eye(4), 1000 * np.eye(2), np.zeros((2, 2)), C.tuning.terminal_ingredient.P, H, umin, umax, terminal="contractive")
    tw = mo.admm_condensed(c, mo.pack_params(x0, xref, uref), mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000))
    assert (res["status"] == tw["status"]).mean() > 0.98 and (res["status"] == 1).mean() > 0.9
    ok = (res["status"] == 1) & (tw["status"] == 1)
    assert (res["iters"][ok] == tw["iters"][ok]).mean() > 0.95 and np.abs(res["u"].reshape(n, -1)[ok] - tw["v"][ok]).max() < 1e-6
    e0 = np.linalg.norm(res["e_x"][:, 0], axis=1); eH = np.linalg.norm(res["e_x"][:, H], axis=1)
    assert (eH[ok] <= np.sqrt(0.9) * e0[ok] + 1e-6).all()
    active = np.abs(eH - np.sqrt(0.9) * e0) < 1e-6
    print(f"contractive H=40 on the shared-memory kernel: ball active on {int(active.sum())} of {n}, solved {int(ok.sum())}, mean iterations {res['iters'].mean():.0f}")
    assert active.sum() >= n // 10


def test_c_abi_error_behaviour(mpc, qt):
    """Errors are return codes + mpcb_last_error (never exceptions across the boundary, never a silent fallback); per-problem
    solver outcomes are data."""
    import ctypes as C
    L = mpc._lib.lib(); lib = mpc._lib
    Cn = make_controller(mpc, qt, 20)
    m = Cn.tuning.modeler
    io = lib.BatchIO(); io.batch = 4
    assert L.mpcb_solve_linear_batch(m._h, C.byref(io)) == -1 and b"x0" in L.mpcb_last_error()          # missing inputs
    x0 = np.zeros((4, 4)); io.x0 = x0.ctypes.data; io.xref = x0.ctypes.data; io.uref = x0.ctypes.data
    io.batch = 0
    assert L.mpcb_solve_linear_batch(m._h, C.byref(io)) == -1 and b"batch" in L.mpcb_last_error()
    io.batch = 4; io.warm_u = x0.ctypes.data                                                                 # warm_u without warm_y
    assert L.mpcb_solve_linear_batch(m._h, C.byref(io)) == -1 and b"warm" in L.mpcb_last_error()
    assert L.mpcb_solve_linear_batch(None, C.byref(io)) == -1
    # design-time failures
    for kw, frag in (({"mpc_terminal_ingredient": "neighborhood"}, "not supported"), ({"mpc_b200_kernel": 9}, "kernel"), ({"mpc_b200_alpha": 2.5}, "settings"),
                     ({"mpc_b200_device": 99}, "device")):
        with pytest.raises(mpc.MpcbError, match=frag):
            make_controller(mpc, qt, 20, **({"terminal": kw.pop("mpc_terminal_ingredient")} if "mpc_terminal_ingredient" in kw else {}), **kw)
    with pytest.raises(mpc.MpcbError, match="on-chip"):
        make_controller(mpc, qt, 50, mpc_b200_kernel=1)                                                      # nz = 100 does not fit the register-resident kernel
    sys_bad = mpc.ConstrainedLinearControlDiscreteSystem(1.5 * np.eye(2), np.zeros((2, 1)), mpc.Hyperrectangle([-1, -1], [1, 1]), mpc.Hyperrectangle([-1], [1]))
    with pytest.raises(mpc.MpcbError):                                                                       # unstabilisable pair: the Riccati iteration must fail loudly
        mpc.proceed_controller(sys_bad, "model_predictive_control", 5, 1, [0, 0], [0], mpc_solver="b200")
    # max_iter is data, not an error
    Cs = make_controller(mpc, qt, 20, mpc_b200_eps_abs=1e-12, mpc_b200_eps_rel=1e-12, mpc_b200_max_iter=10, mpc_b200_check_every=5)
    x0b, xrefb, urefb = qt_batch(qt, 64)
    mpc.update_initialization(Cs, x0b, references=(xrefb, urefb))
    r = mpc.calculate(Cs)
    assert (r["status"] == -2).all() and (r["iters"] == 10).all()


@pytest.mark.parametrize("kernel", [1, 0])
def test_rho_ladder_bounds_the_state_box_tail(mpc, qt, kernel):
    """(kernel 1: the register-resident general-row kernel, forced; 0: the automatic choice, the shared-memory resident general-row kernel at nt = 60.)
    settings.ladder_iter / ladder_kappa: with active state-box rows a batch-wide fixed rho leaves a few problems per 10^4 with
    thousands of iterations (OSQP adapts rho per problem there).  The ladder re-solves what the first pass leaves unsolved with a
    second cached operator (state-box step sizes x kappa).  CUDA vs the twin of the same two-pass scheme: same problems on the
    second rung, same iteration counts, same solutions; and against the single-pass solve: same optima, a bounded tail."""
    H, n, eps = 10, 20000, 1e-7
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    kw = dict(mpc_solver="b200", mpc_state_constraint=True, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000,
              mpc_b200_kernel=kernel)
    C1 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), **kw)
    C2 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_b200_ladder_iter=300, mpc_b200_ladder_kappa=10, **kw)
    assert C1.tuning.modeler.info.kernel == (1 if kernel == 1 else 3)
    rng = np.random.default_rng(7)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    out = []
    for C in (C1, C2):
        mpc.update_initialization(C, x0, references=(xref, qt["u_ref"]))
        out.append({k: v.copy() for k, v in mpc.calculate(C).items()})
    one, two = out
    assert (one["status"] == 1).all() and (two["status"] == 1).all()
    second = one["iters"] > 300
    assert 5 <= second.sum() <= n // 50                                  # a thin tail ...
    assert one["iters"].max() > 3000 and two["iters"].max() < 0.4 * one["iters"].max()     # ... that the second rung cuts
    assert np.array_equal(one["iters"][~second], two["iters"][~second]) and np.array_equal(one["u"][~second], two["u"][~second])
    # the stragglers are the ill-conditioned problems (flat directions of the objective): both solves meet the same residual tolerance,
    # the objective agrees to the parity tolerance, the inputs of the worst straggler to a few 1e-4
    assert mo.u0_metric(one["u"][:, 0], two["u"][:, 0], qt["umin"], qt["umax"]).max() < 5e-4 and np.abs(one["u"] - two["u"]).max() < 5e-3
    assert (np.abs(one["objective"] - two["objective"]) <= 1e-6 * np.abs(one["objective"])).all()
    # the twin of the two-pass scheme
    P = C2.tuning.terminal_ingredient.P
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], xmin, xmax, state_constraint=True)
    sel = np.concatenate([np.flatnonzero(second), np.flatnonzero(~second)[:500]])
    tw = mo.admm_condensed_ladder(c, mo.pack_params(x0[sel], xref[sel], qt["u_ref"]),
                                  mo.AdmmSettings(rho=C2.tuning.modeler.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000), 300, 10.0)
    assert set(tw["second_rung"]) == set(range(int(second.sum())))
    assert (tw["status"] == 1).all() and (two["iters"][sel] == tw["iters"]).mean() > 0.97
    assert np.abs(two["u"][sel].reshape(len(sel), -1) - tw["v"]).max() < 1e-5


def test_streamed_warm_start(mpc, qt):
    """OSQP's warm start (x, y of the previous solve) on the streamed kernel: same second-solve iteration counts and solutions as the
    on-chip kernel (box-only, H = 20) and, with terminal-equality rows (z_g = G x of the warm point), as the on-chip general-row
    kernel (H = 10); the GPU-resident closed loop runs warm-started on a streamed controller."""
    n = 600
    x0, xref, uref = qt_batch(qt, n, seed=23)
    kw = dict(mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=1e-6)
    res = {}
    for kern in (1, 2):
        m = make_controller(mpc, qt, 20, mpc_b200_kernel=kern, **kw).tuning.modeler
        r = m.solve_batch(x0, xref, uref, want=("u", "y"))
        x1 = xref + (x0 - xref) @ qt["A"].T + (r["u"][:, 0] - uref) @ qt["B"].T          # the plants one step later
        res[kern] = m.solve_batch(x1, xref, uref, want=("u", "y"), warm=(r["u"], r["y"]))
        cold = m.solve_batch(x1, xref, uref, want=("u",))
        assert res[kern]["iters"].mean() < cold["iters"].mean()
    a, b = res[1], res[2]
    same = a["iters"] == b["iters"]
    assert same.mean() > 0.99 and (b["status"] == 1).all() and np.abs(a["u"][same] - b["u"][same]).max() < 1e-9 and np.abs(a["u"] - b["u"]).max() < 1e-6
    xr = np.tile(qt["x_ref"], (n, 1)); x0e = xr + 0.002 * np.random.default_rng(3).standard_normal((n, 4))
    out = []
    for kern in (1, 2):
        m = make_controller(mpc, qt, 10, terminal="equality", mpc_b200_kernel=kern, mpc_b200_max_iter=20000, **kw).tuning.modeler
        r = m.solve_batch(x0e, xr, qt["u_ref"], want=("u", "y"))
        out.append(m.solve_batch(x0e * 0.999 + 0.001 * xr, xr, qt["u_ref"], want=("u",), warm=(r["u"], r["y"])))
    ok = (out[0]["status"] == 1) & (out[1]["status"] == 1)
    assert ok.mean() > 0.5 and (out[0]["status"] == out[1]["status"]).mean() > 0.98
    same = ok & (out[0]["iters"] == out[1]["iters"])
    assert same.mean() > 0.45 and np.abs(out[0]["u"][same] - out[1]["u"][same]).max() < 1e-7
    m = make_controller(mpc, qt, 70, mpc_b200_kernel=2, **kw).tuning.modeler
    dev = m.closed_loop(x0[:64], xref[:64], uref, 3, warm_start=True)
    assert (dev["unsolved_steps"] == 0).all()


@pytest.mark.parametrize("kernel", [2, 3])
def test_rho_ladder_on_the_streamed_kernel(mpc, qt, kernel):
    """(kernel 3: the same workload and assertions on the shared-memory resident general-row kernel, which takes it by default since round 2 and
    never synchronises with the host.)
    The same two-rung scheme on the streamed kernel (H = 20 with the state box: nt = 120 general-row problem): first rung capped at 300
    iterations, the unsolved tail re-solved -- through an index map, warm-started in place at the first rung's (x, y) -- with the second cached
    operator.  Against the twin of the two-pass scheme, and against the single-pass solve (same optima, a fraction of the time and tail)."""
    import time
    H, n, eps = 20, 6000, 1e-7
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    kw = dict(mpc_solver="b200", mpc_state_constraint=True, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000,
              mpc_b200_kernel=kernel)
    C1 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), **kw)
    C2 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_b200_ladder_iter=300, mpc_b200_ladder_kappa=10, **kw)
    assert C1.tuning.modeler.info.kernel == kernel and C2.tuning.modeler.info.kernel == kernel and C2.tuning.modeler.info.nt == 120
    rng = np.random.default_rng(7)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    out, ms = [], []
    for C in (C1, C2):
        mpc.update_initialization(C, x0, references=(xref, qt["u_ref"]))
        t0 = time.perf_counter()
        out.append({k: v.copy() for k, v in mpc.calculate(C).items()})
        ms.append((time.perf_counter() - t0) * 1e3)
    one, two = out
    print(f"kernel {kernel} state-box H=20, {n} problems: single pass {ms[0]:.1f} ms (max {one['iters'].max()} iterations), ladder {ms[1]:.1f} ms (max {two['iters'].max()})")
    assert (one["status"] == 1).all() and (two["status"] == 1).all()
    second = one["iters"] > 300
    assert 3 <= second.sum() <= n // 20
    assert two["iters"].max() <= one["iters"].max()
    assert np.array_equal(one["iters"][~second], two["iters"][~second]) and np.array_equal(one["u"][~second], two["u"][~second])
    assert mo.u0_metric(one["u"][:, 0], two["u"][:, 0], qt["umin"], qt["umax"]).max() < 5e-4
    assert (np.abs(one["objective"] - two["objective"]) <= 1e-6 * np.abs(one["objective"])).all()
    P = C2.tuning.terminal_ingredient.P
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], P, H, qt["umin"], qt["umax"], xmin, xmax, state_constraint=True)
    sel = np.concatenate([np.flatnonzero(second), np.flatnonzero(~second)[:300]])
    tw = mo.admm_condensed_ladder(c, mo.pack_params(x0[sel], xref[sel], qt["u_ref"]),
                                  mo.AdmmSettings(rho=C2.tuning.modeler.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000), 300, 10.0)
    assert set(tw["second_rung"]) == set(range(int(second.sum())))
    assert (tw["status"] == 1).all() and (two["iters"][sel] == tw["iters"]).mean() > 0.95
    assert np.abs(two["u"][sel].reshape(len(sel), -1) - tw["v"]).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("H,kernel,terminal,state_box,sigma", [(20, 1, "none", False, 0.0), (20, 1, "none", False, 1e-6), (10, 1, "equality", False, 1e-6),
                                                                 (10, 1, "none", True, 1e-6), (50, 3, "none", False, 0.0), (20, 2, "none", False, 0.0),
                                                                 (10, 2, "equality", False, 1e-6), (10, 2, "none", True, 1e-6), (20, 4, "none", False, 0.0),
                                                                 (100, 4, "none", False, 1e-6), (40, 3, "equality", False, 1e-6), (20, 3, "none", True, 0.0)])
def test_cold_init_parity_on_every_kernel(mpc, qt, H, kernel, terminal, state_box, sigma):
    """settings.cold_init = 1 (kw `mpc_b200_cold_init`): the cold-start point x = clip(Lv p), y_box = -kappa rho (x - Lv p), z_g = G x of every ADMM
    kernel (register-resident with and without general rows, shared-memory, streamed, stage-wise) against the twin's `cold_start_point`: same
    statuses and iteration counts, solutions to round-off; fewer iterations on average than OSQP's zeros, same optima."""
    n, eps = 1300, 1e-7      # more than eight problems per SM: the slot kernels (smaller batches run on the cooperative kernels, which share this code path's arithmetic)
    xmin, xmax = (np.full(4, 0.55), np.full(4, 0.75)) if state_box else (qt["xmin"], qt["xmax"])
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    kw = dict(mpc_solver="b200", mpc_terminal_ingredient=terminal, mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_sigma=sigma,
              mpc_b200_kernel=kernel, mpc_b200_max_iter=20000)
    if state_box: kw["mpc_state_constraint"] = True
    C1 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_b200_cold_init=1, **kw)
    C0 = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), **kw)
    assert C1.tuning.modeler.info.kernel == kernel
    rng = np.random.default_rng(31)
    if state_box: xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    elif terminal == "equality": xref = np.tile(qt["x_ref"], (n, 1)); x0 = xref + rng.uniform(-0.004, 0.004, (n, 4))
    else: x0, xref, _ = qt_batch(qt, n, seed=31)
    uref = qt["u_ref"]
    r1 = C1.tuning.modeler.solve_batch(x0, xref, uref, want=("u", "y", "objective"))
    r0 = C0.tuning.modeler.solve_batch(x0, xref, uref, want=("u", "objective"))
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], C1.tuning.terminal_ingredient.P, H, qt["umin"], qt["umax"], xmin, xmax,
                    state_constraint=state_box, terminal=terminal)
    tw = mo.admm_condensed(c, mo.pack_params(x0, xref, uref), mo.AdmmSettings(rho=C1.tuning.modeler.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, sigma=sigma,
                                                                               max_iter=20000, cold_init=1))
    assert_matches_twin(r1, tw, tight=1e-8, loose=2e-5 if state_box else 1e-6, min_same=0.97, check=5)
    both = (r1["status"] == 1) & (r0["status"] == 1)
    assert both.mean() > 0.99
    assert r1["iters"][both].mean() < r0["iters"][both].mean()
    assert (np.abs(r1["objective"][both] - r0["objective"][both]) <= 1e-6 * np.maximum(1e-3, np.abs(r0["objective"][both]))).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", [1, 0])
def test_rho_ladder_large_second_rung_stays_on_the_slot_kernel(mpc, qt, kernel):
    """A second rung of more than 32 problems per SM is throughput, not latency: the cooperative kernel leaves it to the slot kernel (both are
    enqueued and look at the device-side count).  A first rung capped at 25 iterations sends almost the whole batch there; per-problem results
    equal the twin of the two-pass scheme (a problem's rung does not depend on the others, so the twin runs on a subset)."""
    H, n, eps = 10, 8000, 1e-7
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_state_constraint=True,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000, mpc_b200_kernel=kernel,
                               mpc_b200_ladder_iter=25, mpc_b200_ladder_kappa=10)
    m = C.tuning.modeler
    assert m.info.kernel == (1 if kernel == 1 else 3)
    rng = np.random.default_rng(17)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    res = m.solve_batch(x0, xref, qt["u_ref"], want=("u", "objective"))
    assert (res["status"] == 1).all() and (res["iters"] > 25).sum() > 32 * 148          # the second rung was far too large for the cooperative kernel
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], C.tuning.terminal_ingredient.P, H, qt["umin"], qt["umax"], xmin, xmax, state_constraint=True)
    sel = np.sort(rng.choice(n, 500, replace=False))
    tw = mo.admm_condensed_ladder(c, mo.pack_params(x0[sel], xref[sel], qt["u_ref"]),
                                  mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000), 25, 10.0)
    assert (tw["status"] == 1).all() and (res["iters"][sel] == tw["iters"]).mean() > 0.95
    assert np.abs(res["u"][sel].reshape(len(sel), -1) - tw["v"]).max() < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("H", [10, 20])
def test_rho_ladder_on_a_small_batch(mpc, qt, H):
    """Both rungs of the ladder on the cooperative kernel (a batch of at most eight problems per SM): same second-rung set, iteration counts and
    solutions as the twin of the two-pass scheme."""
    n, eps = 600, 1e-7
    xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200", mpc_state_constraint=True,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=5, mpc_b200_max_iter=20000, mpc_b200_ladder_iter=100, mpc_b200_ladder_kappa=10)
    m = C.tuning.modeler
    rng = np.random.default_rng(41)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    res = m.solve_batch(x0, xref, qt["u_ref"], want=("u", "objective"))
    c = mo.condense(qt["A"], qt["B"], qt["Q"], qt["R"], qt["S"], C.tuning.terminal_ingredient.P, H, qt["umin"], qt["umax"], xmin, xmax, state_constraint=True)
    tw = mo.admm_condensed_ladder(c, mo.pack_params(x0, xref, qt["u_ref"]), mo.AdmmSettings(rho=m.info.rho, eps_abs=eps, eps_rel=eps, check_every=5, max_iter=20000), 100, 10.0)
    assert (res["status"] == 1).all() and (tw["status"] == 1).all()
    assert 5 <= len(tw["second_rung"]) and set(np.flatnonzero(res["iters"] > 100)) == set(tw["second_rung"])
    assert (res["iters"] == tw["iters"]).mean() > 0.97 and np.abs(res["u"].reshape(n, -1) - tw["v"]).max() < 2e-5
