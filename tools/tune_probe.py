import sys, json, numpy as np
sys.path.insert(0, '/root/repo')
import almpc_b200 as mpc, bench
A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
x0, xref, uref = bench.make_batch(16384, 0)
for H in (20, 50, 100, 200):
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7,
                               mpc_b200_check_every=5, mpc_b200_sigma=0.0, mpc_b200_rho_tune=(x0[:1024], xref[:1024], uref, 7))
    t = C.tuning.modeler.rho_tuning
    print(H, [(round(r, 3), round(i, 1)) for r, i in zip(t["candidates"], t["mean_iters"])], "->", round(t["rho"], 3), flush=True)
