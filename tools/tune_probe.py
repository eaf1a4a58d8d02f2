"""Developer probe: what mpcb_tune_rho picks for the quadruple-tank workloads (step-size candidates and their scores)."""
import sys, json, time, pathlib
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import almpc_b200 as mpc, bench

A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
if which == "sweep":
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    x0, xref, uref = bench.make_batch(16384, 0)
    for H in (20, 50, 100, 200):
        C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7,
                                   mpc_b200_check_every=5, mpc_b200_sigma=0.0, mpc_b200_rho_tune=(x0[:4096], xref[:4096], uref, 7))
        t = C.tuning.modeler.rho_tuning
        print(H, [(round(r, 3), round(i, 2)) for r, i in zip(t["candidates"], t["mean_iters"])], "->", round(t["rho"], 3), flush=True)
else:      # state-box rows active on most problems (the workload of test_state_constraint_rows), H = 10 on-chip and H = 20 streamed
    n = 16384
    bx = (np.full(4, 0.55), np.full(4, 0.75))
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(*bx), mpc.Hyperrectangle(umin, umax))
    rng = np.random.default_rng(7)
    xref = rng.uniform(0.70, 0.82, (n, 4)); x0 = rng.uniform(0.62, 0.72, (n, 4))
    for H in (10, 20):
        kw = dict(mpc_solver="b200", mpc_state_constraint=True, mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_max_iter=20000, mpc_b200_sigma=0.0)
        for tune in (None, (x0[:2048], xref[:2048], u_ref, 7)):
            C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_b200_rho_tune=tune, **kw) if tune else \
                mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), **kw)
            m = C.tuning.modeler
            m.solve_batch(x0[:512], xref[:512], u_ref, want=("u0",))
            t0 = time.perf_counter(); r = m.solve_batch(x0, xref, u_ref, want=("u0",)); dt = time.perf_counter() - t0
            print(json.dumps({"H": H, "kernel": m.info.kernel, "tuned": tune is not None, "rho": round(m.info.rho, 4), "wall_ms": round(dt * 1e3, 2), "solve_ms": round(m.timing()["solve_ms"], 2),
                              "mean_iters": round(float(r["iters"].mean()), 1), "max_iters": int(r["iters"].max()), "solved": float((r["status"] == 1).mean()),
                              "tuning": None if tune is None else [(round(a, 3), round(b, 2)) for a, b in zip(m.rho_tuning["candidates"], m.rho_tuning["mean_iters"])]}), flush=True)
