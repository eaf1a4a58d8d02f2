"""Developer micro-bench: solve-kernel time for variants of the QT configs[1] workload (not the driver contract)."""
import argparse, json, sys, pathlib
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import almpc_b200 as mpc
from almpc_b200 import _lib
import bench

def run(H, n, eps, check, sigma, terminal="none", reps=5, full=False, rho=0.0):
    A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_terminal_ingredient=terminal,
                               mpc_b200_eps_abs=eps, mpc_b200_eps_rel=eps, mpc_b200_check_every=check, mpc_b200_sigma=sigma, mpc_b200_rho=rho)
    m = C.tuning.modeler
    x0_h, xref_h, uref_h = bench.make_batch(n, 0)
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
    print(json.dumps({"H": H, "n": n, "eps": eps, "check": check, "sigma": sigma, "terminal": terminal, "full": full, "ms": round(ms, 4),
                      "mean_iters": round(float(it.mean()), 2), "solves_per_s": round(n / ms * 1e3), "tflops": round(fl / ms / 1e9, 2),
                      "frac": round(fl / ms / 1e9 / bench.FP64_PEAK_TFLOPS, 3), "solved": float((status.cpu().numpy() == 1).mean()), "rho": round(m.info.rho, 4)}), flush=True)

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
    elif a.set == "one":
        run(20, 65536, 1e-7, 5, 0.0)
