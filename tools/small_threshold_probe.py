import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
import almpc_b200 as mpc
from almpc_b200 import _lib
import bench
A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
dev = torch.device("cuda", 0)
for H in (10, 20, 50):
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=0.0)
    m = C.tuning.modeler
    for n in (592, 1184, 2368, 4736, 9472):
        x0_h, xref_h, uref_h = bench.make_batch(n, 0)
        io, t = bench._device_io(_lib, dev, n, 4, 2, H, x0_h, xref_h, uref_h)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3): m.solve_batch_device(io, st)
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); m.solve_batch_device(io, st); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(json.dumps({"H": H, "n": n, "per_sm": os.environ.get("MPCB_SMALL_PER_SM", "8"), "us": round(1e3 * float(np.median(ts)), 1)}), flush=True)
    m.close()
