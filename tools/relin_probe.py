import json, pathlib, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import almpc_b200 as mpc
from almpc_b200 import _lib
import bench
A, B, xmin, xmax, umin, umax, x_ref, u_ref, _ = bench.qt_model()
g = json.loads(pathlib.Path('/root/repo/tests/golden/qt_fnn_tanh_model.json').read_text())
f = mpc.Fnn(np.array(g["W_in"]), [(np.array(w), np.array(b)) for w, b in zip(g["W_h"], g["b_h"])], np.array(g["W_out"]), activation=g["activation"])
sys_ = mpc.ConstrainedBlackBoxControlDiscreteSystem(f, 4, 2, mpc.Hyperrectangle(xmin, xmax), mpc.Hyperrectangle(umin, umax))
C = mpc.proceed_controller(sys_, "model_predictive_control", 20, 5, list(x_ref), list(u_ref), mpc_solver="b200", mpc_programming_type="non_linear")
mod = C.tuning.modeler
dev = torch.device("cuda", 0)
def setup(n, outputs=True, seed=0):
    x0_h, xref_h, _ = bench.make_batch(n, seed)
    uref_h = np.random.default_rng(3).uniform(0.8, 2.2, (n, 2))
    t = dict(x0=torch.from_numpy(x0_h).to(dev), xref=torch.from_numpy(xref_h).to(dev), uref=torch.from_numpy(uref_h).to(dev),
             status=torch.empty(n, dtype=torch.int32, device=dev), iters=torch.empty(n, dtype=torch.int32, device=dev), inner=torch.empty(n, dtype=torch.int32, device=dev),
             u=torch.empty((n, 20, 2), dtype=torch.float64, device=dev), x=torch.empty((n, 21, 4), dtype=torch.float64, device=dev))
    io = _lib.BatchIO(); io.batch = n; io.x0 = t['x0'].data_ptr(); io.xref = t['xref'].data_ptr(); io.uref = t['uref'].data_ptr()
    io.status = t['status'].data_ptr(); io.iters = t['iters'].data_ptr(); io.inner_iters = t['inner'].data_ptr()
    if outputs: io.u = t['u'].data_ptr(); io.x = t['x'].data_ptr()
    return io, t
st = torch.cuda.current_stream().cuda_stream
def timeit(io, k=1):
    mod.solve_batch_device(io, st, method="linear"); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k): mod.solve_batch_device(io, st, method="linear")
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)
io16, t16 = setup(16384); io64, t64 = setup(65536); io64n, t64n = setup(65536, outputs=False)
print("16384 x1", timeit(io16), " x8 back-to-back", timeit(io16, 8))
print("65536 x1", timeit(io64), " no u/x outputs", timeit(io64n))
inner = t64['inner'].cpu().numpy(); print("inner iters: mean", inner.mean(), "max", inner.max(), "p99", np.percentile(inner, 99), "first16k mean", inner[:16384].mean(), "last16k mean", inner[-16384:].mean())
# the first 16384 problems of the 65536 batch alone
io_sub = _lib.BatchIO(); io_sub.batch = 16384; io_sub.x0 = t64['x0'].data_ptr(); io_sub.xref = t64['xref'].data_ptr(); io_sub.uref = t64['uref'].data_ptr(); io_sub.status = t64['status'].data_ptr(); io_sub.iters = t64['iters'].data_ptr(); io_sub.inner_iters = t64['inner'].data_ptr()
print("first 16384 of the 65536 batch", timeit(io_sub))
