"""One handle, several GPUs, one process (mpcb_settings.n_devices / device_ids; kw `mpc_b200_devices`): the batch is cut into contiguous
shards, one per device, with no exchange during the solve -- the per-problem results must be BIT-identical to the single-device
handle's, through the host entry (one host thread per device), the device entry (NVLink peer copies ordered by events) and the
closed-loop entry.  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import numpy as np
import pytest

from conftest import load_nn_fixture, qt_batch
from test_gpu_linear import make_controller

pytestmark = pytest.mark.gpu


def n_gpus(mpc):
    return mpc._lib.lib().mpcb_device_count()


@pytest.mark.parametrize("H,kernel", [(20, 0), (50, 0), (80, 0), (30, 2)])
def test_multi_device_host_and_device_entries_equal_single_device(mpc, qt, H, kernel):
    if n_gpus(mpc) < 2: pytest.skip("needs >= 2 GPUs")
    import torch
    devs = list(range(min(n_gpus(mpc), 4)))
    n = 5001                                             # ragged shards
    kw = dict(mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=0.0, mpc_b200_kernel=kernel)
    one = make_controller(mpc, qt, H, **kw).tuning.modeler
    many = make_controller(mpc, qt, H, mpc_b200_devices=devs, **kw).tuning.modeler
    assert many.info.kernel == one.info.kernel and many.info.device == devs[0]
    x0, xref, uref = qt_batch(qt, n, seed=41)
    want = ("u", "e_u", "x", "e_x", "u0", "objective", "y")
    a = one.solve_batch(x0, xref, uref, want=want)
    b = many.solve_batch(x0, xref, uref, want=want)
    for k in want + ("status", "iters", "prim_res", "dual_res"):
        assert np.array_equal(a[k], b[k]), k
    assert many.timing()["kernel_launches"] >= len(devs) * (one.timing()["kernel_launches"] if kernel != 2 else 1)
    # warm start through the sharded host entry
    w1 = one.solve_batch(x0, xref, uref, want=("u",), warm=(a["u"], a["y"])) if one.info.kernel != 2 else None
    if w1 is not None:
        w2 = many.solve_batch(x0, xref, uref, want=("u",), warm=(a["u"], a["y"]))
        assert np.array_equal(w1["u"], w2["u"]) and np.array_equal(w1["iters"], w2["iters"])
    # device entry: buffers on device_ids[0], shards fanned out / gathered over NVLink peer copies
    dev = torch.device("cuda", devs[0])
    t = {"x0": torch.from_numpy(x0).to(dev), "xref": torch.from_numpy(xref).to(dev), "uref": torch.from_numpy(uref).to(dev),
         "u": torch.empty((n, H, 2), dtype=torch.float64, device=dev), "x": torch.empty((n, H + 1, 4), dtype=torch.float64, device=dev),
         "u0": torch.empty((n, 2), dtype=torch.float64, device=dev), "objective": torch.empty(n, dtype=torch.float64, device=dev),
         "status": torch.empty(n, dtype=torch.int32, device=dev), "iters": torch.empty(n, dtype=torch.int32, device=dev)}
    io = mpc._lib.BatchIO(); io.batch = n; io.uref_broadcast = 1
    for k, v in t.items(): setattr(io, k, v.data_ptr())
    with torch.cuda.device(dev):
        for _ in range(2):                               # twice: the second call reuses every peer buffer and event
            for k in ("u", "x", "u0", "objective"): t[k].zero_()
            many.solve_batch_device(io, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            for k in ("u", "x", "u0", "objective", "status", "iters"):
                assert np.array_equal(t[k].cpu().numpy(), a[k]), k
    # small batches stay on the root device (the closed-loop, one-problem-at-a-time use)
    s1 = many.solve_batch(x0[:7], xref[:7], uref)
    assert np.array_equal(s1["u"], a["u"][:7])


def test_multi_device_closed_loop_and_errors(mpc, qt):
    if n_gpus(mpc) < 2: pytest.skip("needs >= 2 GPUs")
    kw = dict(mpc_b200_eps_abs=1e-7, mpc_b200_eps_rel=1e-7, mpc_b200_check_every=5, mpc_b200_sigma=0.0)
    one = make_controller(mpc, qt, 20, **kw).tuning.modeler
    many = make_controller(mpc, qt, 20, mpc_b200_devices=[0, 1], **kw).tuning.modeler
    x0, xref, uref = qt_batch(qt, 3000, seed=42)
    a = one.closed_loop(x0, xref, uref, 5); b = many.closed_loop(x0, xref, uref, 5)
    for k in a: assert np.array_equal(a[k], b[k]), k
    with pytest.raises(mpc.MpcbError, match="duplicate"):
        make_controller(mpc, qt, 20, mpc_b200_devices=[0, 0])
    with pytest.raises(mpc.MpcbError, match="out of range"):
        make_controller(mpc, qt, 20, mpc_b200_devices=[0, 99])


def test_multi_device_nmpc_host_entry(mpc, qt):
    if n_gpus(mpc) < 2: pytest.skip("needs >= 2 GPUs")
    from test_gpu_nmpc import make_system, scenario
    m = load_nn_fixture("qt_resnet_model.json")
    x0, xref, uref = scenario(qt, 2049)
    out = []
    for devs in (None, [0, 1]):
        kw = {} if devs is None else {"mpc_b200_devices": devs}
        C = mpc.proceed_controller(make_system(mpc, qt, m), "model_predictive_control", 20, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200",
                                   mpc_programming_type="non_linear", **kw)
        out.append(C.tuning.modeler.solve_batch(x0, xref, uref))
    for k in ("u", "x", "objective", "status", "iters", "inner_iters"):
        assert np.array_equal(out[0][k], out[1][k]), k
