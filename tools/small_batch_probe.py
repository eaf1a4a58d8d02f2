"""Latency of small batches on controllers WITH general rows (terminal equality / state box), host-array entry, pinned zero-copy path:
the CTA-cooperative kernel (default for batches of up to 8 problems per SM) against the slot kernels (MPCB_NO_SMALL_COOP=1).
Run twice: `python tools/small_batch_probe.py` and `MPCB_NO_SMALL_COOP=1 python tools/small_batch_probe.py`."""
import json, os, sys, time, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import almpc_b200 as mpc
import bench

A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
rng = np.random.default_rng(0)
for label, H, extra in (("terminal equality H=20 (nt=44)", 20, dict(mpc_terminal_ingredient="equality")), ("terminal equality H=40 (nt=84)", 40, dict(mpc_terminal_ingredient="equality")),
                        ("state box H=10 (nt=60)", 10, dict(mpc_state_constraint=True)), ("state box H=20 (nt=120)", 20, dict(mpc_state_constraint=True))):
    sb = "mpc_state_constraint" in extra
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(np.full(4, 0.55) if sb else xmin, np.full(4, 0.75) if sb else xmax), mpc.Hyperrectangle(umin, umax))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7,
                               mpc_b200_check_every=5, mpc_b200_sigma=0.0, mpc_b200_max_iter=20000, **extra)
    m = C.tuning.modeler
    for n in (1, 8, 64, 512):
        if sb: x0 = rng.uniform(0.62, 0.72, (n, 4)); xr = rng.uniform(0.70, 0.82, (n, 4))
        else: xr = np.tile(x_ref, (n, 1)); x0 = xr + 0.0015 * rng.standard_normal((n, 4))
        r = m.solve_batch(x0, xr, np.asarray(u_ref), want=("u0",))
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); r = m.solve_batch(x0, xr, np.asarray(u_ref), want=("u0",)); ts.append((time.perf_counter() - t0) * 1e6)
        print(json.dumps({"workload": label, "kernel": m.info.kernel, "batch": n, "cold_start_p50_us": round(float(np.median(ts)), 1), "mean_iters": float(r["iters"].mean()),
                          "max_iters": int(r["iters"].max()), "no_small_coop": bool(os.environ.get("MPCB_NO_SMALL_COOP"))}), flush=True)
    m.close()
