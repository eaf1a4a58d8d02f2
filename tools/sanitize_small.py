"""Small invocations of the newer kernels for compute-sanitizer (memcheck / racecheck): batched DARE, re-linearised solve
(plain / state box / terminal equality / contractive), SQP with the contractive ball, NMPC closed loop, linear on-chip kernels."""
import json, pathlib, sys
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import almpc_b200 as mpc
import bench

A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
g = json.loads((ROOT / "tests" / "golden" / "qt_fnn_tanh_model.json").read_text())
f = mpc.Fnn(np.array(g["W_in"]), [(np.array(w), np.array(b)) for w, b in zip(g["W_h"], g["b_h"])], np.array(g["W_out"]), activation=g["activation"])
rng = np.random.default_rng(0)
n = 37
x0 = rng.uniform(0.5, 0.8, (n, 4)); xref = rng.uniform(0.55, 0.75, (n, 4)); uref = rng.uniform(1.0, 2.0, (n, 2))
Q, R, S = 100 * np.eye(4), 0.1 * np.eye(2), np.zeros((2, 2))
_, Aj, Bj = f.jacobian(xref, uref)
P, st = mpc.dare_batch(Aj, Bj, Q, R); assert (st > 0).all()
for terminal, sc in (("none", False), ("none", True), ("equality", False), ("contractive", False)):
    mod = mpc.B200NonlinearModeler(f, Q, R, S, None, umin, umax, np.full(4, 0.5), np.full(4, 0.85), 6, x_ref, u_ref, state_constraint=sc, terminal=terminal)
    r1 = mod.solve_batch(x0, xref, uref, want=("u", "x", "objective", "y"), method="linear")
    r2 = mod.solve_batch(x0, xref, uref, want=("u", "x", "objective", "y"))
    cl = mod.closed_loop(x0, xref, uref, 3, warm_start=True)
    print(terminal, sc, np.unique(r1["status"]), np.unique(r2["status"]), cl["unsolved_steps"].sum())
    mod.close()
sys_ = mpc.ConstrainedLinearControlDiscreteSystem(A, B, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
for H in (20, 30):
    C = mpc.proceed_controller(sys_, "model_predictive_control", H, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5,
                               mpc_b200_sigma=0.0)
    mpc.update_initialization(C, x0, references=(xref, u_ref)); r = mpc.calculate(C)
    print("linear H", H, np.unique(r["status"]), C.tuning.modeler.info.kernel)
print("sanitize_small ok")
