"""Developer micro-bench: solve-kernel time for variants of the QT configs[1] workload (not the driver contract)."""
import argparse, json, sys, pathlib
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import almpc_b200 as mpc
from almpc_b200 import _lib
import bench

def run(H, n, eps, check, sigma, terminal="none", reps=5, full=False, rho=0.0, near=0.0, state_box=False, Qw=None, Rw=None, max_iter=20000, ladder=0, kernel=0, cold_init=0):
    """near > 0: x0 = x_ref + near * N(0, I) with the design reference (feasible terminal constraints); state_box: tight box + references beyond it"""
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
    if state_box: xmin, xmax = np.full(4, 0.55), np.full(4, 0.75)
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    extra = {}
    if state_box: extra["mpc_state_constraint"] = True
    if Qw is not None: extra["mpc_Q"] = Qw
    if Rw is not None: extra["mpc_R"] = Rw
    if ladder: extra["mpc_b200_ladder_iter"] = ladder
    extra["mpc_b200_cold_init"] = cold_init
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_terminal_ingredient=terminal,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=sigma, mpc_b200_rho=rho,
                               mpc_b200_max_iter=max_iter, mpc_b200_kernel=kernel, **extra)
    m = C.tuning.modeler
    x0_h, xref_h, uref_h = bench.make_batch(n, 0)
    rng = np.random.default_rng(7)
    if near > 0: xref_h = np.tile(x_ref, (n, 1)); x0_h = xref_h + near * rng.standard_normal((n, 4))
    if state_box: xref_h = rng.uniform(0.70, 0.82, (n, 4)); x0_h = rng.uniform(0.62, 0.72, (n, 4))
    dev = torch.device("cuda", 0)
    x0 = torch.from_numpy(x0_h).to(dev); xref = torch.from_numpy(xref_h).to(dev); uref = torch.from_numpy(uref_h).to(dev)
    status = torch.empty(n, dtype=torch.int32, device=dev); iters = torch.empty(n, dtype=torch.int32, device=dev)
    io = _lib.BatchIO(); io.batch = n; io.x0 = x0.data_ptr(); io.xref = xref.data_ptr(); io.uref = uref.data_ptr(); io.uref_broadcast = 1
    io.status = status.data_ptr(); io.iters = iters.data_ptr()
    outs = []
    if full:
        u = torch.empty((n, H, 2), dtype=torch.float64, device=dev); eu = torch.empty_like(u)
        x = torch.empty((n, H + 1, 4), dtype=torch.float64, device=dev); ex = torch.empty_like(x)
        io.u = u.data_ptr(); io.e_u = eu.data_ptr(); io.x = x.data_ptr(); io.e_x = ex.data_ptr(); outs = [u, eu, x, ex]
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): m.solve_batch_device(io, st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); m.solve_batch_device(io, st); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    it = iters.cpu().numpy()
    fl = bench.algorithmic_flops(m.info, it, check)
    ms = min(ts)
    extra_out = {}
    if m.info.kernel == 4:      # stage-wise kernel: its own arithmetic (68 multiply-adds per stage and iteration for nx=4, nu=2) and its state traffic
        nx, nu = m.info.nx, m.info.nu
        sfl = float(it.astype(float).sum()) * H * 2 * (2 * nx * nx + 4 * nx * nu + nu * nu)
        sby = float(it.astype(float).sum()) * (6 if sigma == 0 else 9) * m.info.nz * 8
        extra_out = {"stage_tflops": round(sfl / ms / 1e9, 2), "state_GBps": round(sby / ms / 1e6, 1)}
    print(json.dumps({**extra_out, "H": H, "n": n, "eps": eps, "check": check, "sigma": sigma, "terminal": terminal, "state_box": state_box, "kernel": m.info.kernel, "nt": m.info.nt,
                      "full": full, "ladder": ladder, "cold_init": cold_init, "ms": round(ms, 4), "max_iters": int(it.max()),
                      "mean_iters": round(float(it.mean()), 2), "solves_per_s": round(n / ms * 1e3), "tflops": round(fl / ms / 1e9, 2),
                      "frac": round(fl / ms / 1e9 / bench.FP64_PEAK_TFLOPS, 3), "solved": float((status.cpu().numpy() == 1).mean()), "rho": round(m.info.rho, 4)}), flush=True)

def lti64(seed=1, nx=64, nu=16):
    """random stable LTI."""
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((nx, nx)); A = 0.95 * G / np.abs(np.linalg.eigvals(G)).max(); B = rng.standard_normal((nx, nu)) / 8
    return A, B


def run_lti(H=50, n=8192, eps=1e-5, check=10, sigma=0.0, scale=0.3, terminal="equality", reps=2, rho=0.0, max_iter=4000, full=False):
    import time
    A, B = lti64()
    nx, nu = B.shape
    big = 1e3
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(-big * np.ones(nx), big * np.ones(nx)), mpc.Hyperrectangle(-np.ones(nu), np.ones(nu)))
    t0 = time.perf_counter()
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 1, [0.0] * nx, [0.0] * nu, mpc_solver="b200", mpc_terminal_ingredient=terminal,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=sigma, mpc_b200_rho=rho, mpc_b200_max_iter=max_iter)
    t_design = time.perf_counter() - t0
    m = C.tuning.modeler
    rng = np.random.default_rng(3)
    x0_h = scale * rng.standard_normal((n, nx)); xref_h = np.zeros(nx); uref_h = np.zeros(nu)
    dev = torch.device("cuda", 0)
    x0 = torch.from_numpy(x0_h).to(dev); xref = torch.from_numpy(xref_h).to(dev); uref = torch.from_numpy(uref_h).to(dev)
    status = torch.empty(n, dtype=torch.int32, device=dev); iters = torch.empty(n, dtype=torch.int32, device=dev)
    u0 = torch.empty((n, nu), dtype=torch.float64, device=dev)
    io = _lib.BatchIO(); io.batch = n; io.x0 = x0.data_ptr(); io.xref = xref.data_ptr(); io.uref = uref.data_ptr(); io.uref_broadcast = 1; io.xref_broadcast = 1
    io.status = status.data_ptr(); io.iters = iters.data_ptr(); io.u0 = u0.data_ptr()
    if full:
        uu = torch.empty((n, H, nu), dtype=torch.float64, device=dev); eu = torch.empty_like(uu)
        xx = torch.empty((n, H + 1, nx), dtype=torch.float64, device=dev); ex = torch.empty_like(xx); ob = torch.empty(n, dtype=torch.float64, device=dev)
        io.u = uu.data_ptr(); io.e_u = eu.data_ptr(); io.x = xx.data_ptr(); io.e_x = ex.data_ptr(); io.objective = ob.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    m.solve_batch_device(io, st); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); m.solve_batch_device(io, st); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    it = iters.cpu().numpy(); stt = status.cpu().numpy()
    i = m.info
    fl = float((it.astype(float) * 2 * i.nt * i.nt).sum())
    ms = min(ts)
    print(json.dumps({"cfg": "lti", "nx": nx, "nu": nu, "H": H, "n": n, "eps": eps, "check": check, "scale": scale, "terminal": terminal, "nt": i.nt, "nt_pad": i.nt_pad,
                      "kernel": i.kernel, "rho": round(i.rho, 4), "design_s": round(t_design, 2), "ms": round(ms, 2), "mean_iters": round(float(it.mean()), 1),
                      "max_iters": int(it.max()), "solved": float((stt == 1).mean()), "infeasible": float((stt == -3).mean()), "maxiter": float((stt == -2).mean()),
                      "solves_per_s": round(n / ms * 1e3), "tflops_active_rows": round(fl / ms / 1e9, 2), "launches": m.timing()["kernel_launches"]}), flush=True)


def run_nmpc(fixture="qt_resnet_model.json", H=20, n=4096, reps=3, method="non_linear", **kw):
    """NMPC with a neural dynamics model, SQP kernel."""
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
    g = json.loads((ROOT / "tests" / "golden" / fixture).read_text())          # the fixture weights, read directly (no oracle import here)
    cls = {"fnn": mpc.Fnn, "resnet": mpc.ResNet, "polynet": mpc.PolyNet, "densenet": mpc.DenseNet}[g["arch"]]
    f = cls(np.array(g["W_in"]), [(np.array(w), np.array(b)) for w, b in zip(g["W_h"], g["b_h"])], np.array(g["W_out"]), activation=g["activation"])
    sys_ = mpc.ConstrainedBlackBoxControlDiscreteSystem(f, 4, 2, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_programming_type="non_linear", **kw)
    mod = C.tuning.modeler
    x0_h, xref_h, uref_h = bench.make_batch(n, 0)
    dev = torch.device("cuda", 0)
    if method == "linear":          # per-problem references: every problem is its own linearisation, Riccati equation and QP
        rng = np.random.default_rng(3)
        uref_h = rng.uniform(0.8, 2.2, (n, 2))
    x0 = torch.from_numpy(x0_h).to(dev); xref = torch.from_numpy(xref_h).to(dev); uref = torch.from_numpy(uref_h).to(dev)
    status = torch.empty(n, dtype=torch.int32, device=dev); iters = torch.empty(n, dtype=torch.int32, device=dev); inner = torch.empty(n, dtype=torch.int32, device=dev)
    u = torch.empty((n, H, 2), dtype=torch.float64, device=dev); x = torch.empty((n, H + 1, 4), dtype=torch.float64, device=dev)
    eu = torch.empty_like(u); ex = torch.empty_like(x); obj = torch.empty(n, dtype=torch.float64, device=dev)
    io = _lib.BatchIO(); io.batch = n; io.x0 = x0.data_ptr(); io.xref = xref.data_ptr(); io.uref = uref.data_ptr(); io.uref_broadcast = 0 if method == "linear" else 1
    io.status = status.data_ptr(); io.iters = iters.data_ptr(); io.inner_iters = inner.data_ptr()
    io.u = u.data_ptr(); io.x = x.data_ptr(); io.e_u = eu.data_ptr(); io.e_x = ex.data_ptr(); io.objective = obj.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    mod.solve_batch_device(io, st, method=method); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); mod.solve_batch_device(io, st, method=method); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    it = iters.cpu().numpy(); inn = inner.cpu().numpy(); stt = status.cpu().numpy()
    ms = min(ts)
    L = _lib.lib()
    if hasattr(L, "mpcb_debug_nmpc_prof"):      # development build with -DMPCB_NMPC_PROF
        import ctypes
        buf = (ctypes.c_ulonglong * 8)()
        L.mpcb_debug_nmpc_prof(buf, 1)
        tot = sum(buf) or 1
        names = ["linearise-other", "q+inverse", "admm", "final accept", "line search/init", "nn_eval+jac", "gamma+Wgamma", "K,g accumulate"]
        print(json.dumps({"nmpc_phase_share": {nm: round(buf[i] / tot, 3) for i, nm in enumerate(names)}, "cycles_per_problem_per_launch": round(tot / n / (reps + 1))}), flush=True)
    print(json.dumps({"cfg": "nmpc" if method == "non_linear" else "relinearized", "fixture": fixture, "H": H, "n": n, "ms": round(ms, 3), "solves_per_s": round(n / ms * 1e3), "sqp_iters_mean": round(float(it.mean()), 2),
                      "sqp_iters_max": int(it.max()), "inner_mean": round(float(inn.mean()), 1), "solved": float((stt == 1).mean()), "stalled": float((stt == 2).mean()),
                      "maxiter": float((stt == -2).mean()), "rho": round(mod.design()["rho"], 4)}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--set", default="qt")
    a = ap.parse_args()
    if a.set == "qt":
        for sigma in (1e-6, 0.0):
            for check in (5, 10):
                run(20, 65536, 1e-7, check, sigma)
        run(20, 65536, 1e-7, 5, 0.0, full=True)
        run(20, 65536, 1e-3, 25, 1e-6)
        run(20, 65536, 1e-3, 5, 0.0)
    elif a.set == "steady":      # no drain, no spread: every problem runs exactly 50 iterations and the batch is a whole number of slot loads
        for n in (14208 * 4, 14208 * 16, 65536):
            run(20, n, 1e-300, 5, 0.0, max_iter=50)
        run(20, 65536 * 4, 1e-7, 5, 0.0)
    elif a.set == "coldinit":    # A/B of settings.cold_init: steady state (every problem exactly 50 iterations: isolates the refill cost), the headline batch, the sweep
        for ci in (0, 1):
            run(20, 14208 * 16, 1e-300, 5, 0.0, max_iter=50, cold_init=ci)
            run(20, 65536, 1e-7, 5, 0.0, cold_init=ci)
            run(20, 65536, 1e-7, 5, 0.0, full=True, cold_init=ci)
        for H in (10, 30, 50, 100, 200):
            for ci in (0, 1): run(H, 16384, 1e-7, 5, 0.0, reps=3, cold_init=ci)
        for ci in (0, 1): run(10, 65536, 1e-7, 5, 0.0, reps=3, state_box=True, cold_init=ci, ladder=300)
    elif a.set == "sbox20":      # state box at H = 20 (nt = 120): shared-memory resident general-row kernel (auto) vs the streamed kernel, with and without the rho ladder
        for n in (6000, 16384):
            for kern in (0, 2):
                for lad in (0, 300):
                    if kern == 2 and n > 6000 and lad: continue
                    run(20, n, 1e-7, 5, 0.0, state_box=True, kernel=kern, ladder=lad, reps=2)
        run(40, 16384, 1e-7, 10, 0.0, terminal="equality", near=0.02, kernel=0, reps=3)
        run(40, 16384, 1e-7, 10, 0.0, terminal="equality", near=0.02, kernel=2, reps=3)
    elif a.set == "sbox10":      # state box at H = 10 (nt = 60): register-resident general-row kernel vs the shared-memory resident one (forced)
        for kern in (1, 3):
            for lad in (0, 300):
                run(10, 65536, 1e-7, 5, 0.0, state_box=True, kernel=kern, ladder=lad, reps=2)
        for kern in (1, 3):
            run(26, 65536, 1e-7, 5, 0.0, terminal="equality", near=0.02, kernel=kern, reps=3)      # nt = 56
            run(30, 65536, 1e-7, 5, 0.0, terminal="equality", near=0.02, kernel=kern, reps=3)      # nt = 64
    elif a.set == "eqab":        # terminal equality, nt = 32 .. 64: register-resident general-row kernel vs the shared-memory resident one (forced)
        for H in (14, 18, 20, 22):
            for kern in (1, 3):
                run(H, 65536, 1e-7, 5, 0.0, terminal="equality", near=0.02, kernel=kern, reps=3)
        for kern in (1, 3): run(5, 65536, 1e-7, 5, 0.0, state_box=True, kernel=kern, reps=3)      # state box H = 5: nt = 30
    elif a.set == "eqab2":       # the same A/B on FEASIBLE batches (x0 within 0.004 of the reference): throughput, not the iteration-cap tail
        for H in (16, 20, 22, 26, 30):
            for kern in (1, 3):
                run(H, 65536, 1e-7, 5, 0.0, terminal="equality", near=0.0015, kernel=kern, reps=3)
    elif a.set == "eq40one":     # one throughput-bound launch of the shared-memory resident general-row kernel (ncu target)
        run(40, 65536, 1e-7, 10, 0.0, terminal="equality", near=0.0015, kernel=0, reps=2)
    elif a.set == "coopone":     # one state-box solve with the ladder: first rung on the shared-memory general-row kernel, second on the cooperative kernel (ncu target)
        run(20, 16384, 1e-7, 5, 0.0, state_box=True, kernel=0, ladder=300, reps=1)
    elif a.set == "steady1":
        run(20, 14208 * 8, 1e-300, 5, 0.0, max_iter=50, reps=3)
    elif a.set == "one":
        run(20, 65536, 1e-7, 5, 0.0)
    elif a.set == "onefull":
        run(20, 65536, 1e-7, 5, 0.0, full=True, reps=2)
    elif a.set == "hsweep":
        for H in (10, 20, 30, 50, 75, 100, 150, 200):
            run(H, 16384, 1e-7, 5, 0.0, reps=2)
    elif a.set == "ricsweep":    # config 4 with the stage-wise kernel next to the condensed ones: the crossover study
        for H in (10, 20, 30, 50, 75, 100, 150, 200):
            run(H, 16384, 1e-7, 5, 0.0, reps=3, kernel=4)
            run(H, 16384, 1e-7, 5, 0.0, reps=3, kernel=0)
        for H in (20, 50):
            run(H, 65536, 1e-7, 5, 0.0, reps=3, kernel=4)
    elif a.set == "ric1":
        run(50, 16384, 1e-7, 5, 0.0, reps=2, kernel=4)
    elif a.set == "ric100":
        run(100, 16384, 1e-7, 5, 0.0, reps=1, kernel=4)
    elif a.set == "ric200":
        run(200, 16384, 1e-7, 5, 0.0, reps=1, kernel=4)
    elif a.set == "ric20big":
        run(20, 65536, 1e-7, 5, 0.0, reps=1, kernel=4)
    elif a.set == "rows":        # general-row variants of the QT controller (a3 / a4 rows of SURVEY 8a)
        run(20, 65536, 1e-7, 5, 0.0, terminal="equality", near=0.002, reps=3)
        run(10, 65536, 1e-7, 5, 0.0, state_box=True, reps=3)
        run(10, 65536, 1e-7, 5, 0.0, state_box=True, reps=3, ladder=300)      # rho ladder: second rung for the tail
        run(20, 16384, 1e-7, 5, 0.0, state_box=True, reps=2)
        run(3, 65536, 1e-8, 5, 0.0, terminal="contractive", near=0.15, Qw=1.0, Rw=10.0, reps=3)
    elif a.set == "nmpc":
        for fx in ("qt_resnet_model.json", "qt_fnn_tanh_model.json", "qt_resnet_swish_model.json", "qt_polynet_tanh_model.json", "qt_densenet_tanh_model.json",
                   "qt_fnn_model.json"):
            run_nmpc(fx)
        run_nmpc("qt_resnet_model.json", n=65536)
    elif a.set == "relin":       # per-problem linearisation + DARE + condensed QP in one launch
        for fx, H in (("qt_fnn_tanh_model.json", 20), ("qt_resnet_model.json", 20), ("qt_fnn_tanh_model.json", 10), ("qt_densenet_tanh_model.json", 20)):
            run_nmpc(fx, H=H, n=65536, method="linear")
        run_nmpc("qt_fnn_tanh_model.json", H=20, n=65536, method="linear", mpc_b200_eps_abs=1e-7)
    elif a.set == "relinscale":
        for n in (4096, 16384, 32768, 65536, 131072):
            run_nmpc("qt_fnn_tanh_model.json", H=20, n=n, method="linear", reps=3)
    elif a.set == "relucap":     # the reference's relu FNN fixture: how much of its time is the inner iteration cap of non-converging QPs
        for cap in (4000, 1000, 500, 200):
            run_nmpc("qt_fnn_model.json", mpc_b200_max_iter=cap)
        for cap in (4000, 500, 200):
            run_nmpc("qt_fnn_tanh_model.json", mpc_b200_max_iter=cap)
    elif a.set == "smemk":       # shared-memory kernel range (64 < nt <= 120)
        for H in (36, 40, 44, 48, 50, 60):
            run(H, 16384, 1e-7, 5, 0.0, reps=3)
    elif a.set == "small":       # small batches: about one problem per slot at full occupancy
        for H, n in ((10, 16384), (20, 16384), (30, 16384), (20, 32768), (30, 32768), (20, 8192)):
            run(H, n, 1e-7, 5, 0.0, reps=3)
    elif a.set == "h24sig":
        run(24, 65536, 1e-7, 5, 1e-6, reps=3)
    elif a.set == "h24":
        for H in (24, 28, 32):
            run(H, 65536, 1e-7, 5, 0.0, reps=3)
    elif a.set == "relin1":
        run_nmpc("qt_fnn_tanh_model.json", H=20, n=16384, method="linear", reps=1)
    elif a.set == "nmpc1":
        run_nmpc("qt_fnn_tanh_model.json", reps=1)
    elif a.set == "lti_tuned":   # config 3 as bench.py runs it: eps 1e-7, the tuned step size, full outputs
        run_lti(scale=1.0, eps=1e-7, rho=296.0242231923552, reps=2, full=True)
    elif a.set == "lti_check":   # config 3 at the tuned step size: check period 10 (the parity setting) vs 5 now that check GEMMs skip blocks that cannot terminate
        for chk in (10, 5):
            run_lti(scale=1.0, eps=1e-7, rho=296.0242231923552, reps=3, full=True, check=chk)
    elif a.set == "lti1":
        run_lti(scale=1.0, reps=1)
    elif a.set == "h50":
        run(50, 16384, 1e-7, 5, 0.0, reps=1)
    elif a.set == "lti":
        for scale in (1.0, 3.0, 10.0, 30.0):
            run_lti(scale=scale)
        run_lti(scale=0.3, terminal="none")
