"""ctypes binding of libmpcb200.so (include/mpcb200.h).  The product path: if the CUDA library is missing or no
B200 is visible, calls FAIL LOUDLY -- there is no CPU fallback (the oracle under oracle/ is test infrastructure and is
never imported from here)."""
from __future__ import annotations

import ctypes as C
import os
import pathlib

_PKG = pathlib.Path(__file__).resolve().parent
LIB_PATH = pathlib.Path(os.environ["MPCB200_LIB"]) if os.environ.get("MPCB200_LIB") else _PKG / "libmpcb200.so"   # the override serves A/B builds of the kernels (tools/)

MPCB_OK = 0
KERNEL_AUTO, KERNEL_ONCHIP, KERNEL_STREAMED, KERNEL_ONCHIP_SMEM, KERNEL_RICCATI = 0, 1, 2, 3, 4
TERMINAL_NONE, TERMINAL_EQUALITY, TERMINAL_CONTRACTIVE = 0, 1, 2
STATUS_SOLVED, STATUS_SOLVED_INACCURATE, STATUS_MAX_ITER, STATUS_PRIMAL_INFEASIBLE, STATUS_DESIGN_FAILED = 1, 2, -2, -3, -20
NN_FNN, NN_RESNET, NN_POLYNET, NN_DENSENET = 0, 1, 2, 3
ACTIVATION_IDS = {"relu": 0, "tanh": 1, "sigmoid": 2, "swish": 3, "identity": 4}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class Settings(C.Structure):
    _fields_ = [("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_prim_inf", C.c_double), ("rho", C.c_double),
                ("rho_eq_scale", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double), ("max_iter", C.c_int32),
                ("check_every", C.c_int32), ("device", C.c_int32), ("kernel", C.c_int32), ("ladder_iter", C.c_int32), ("ladder_kappa", C.c_int32),
                ("n_devices", C.c_int32), ("device_ids", C.c_int32 * 8), ("cold_init", C.c_int32)]


class LinearDesc(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("horizon", C.c_int32), ("A", _dp), ("B", _dp), ("Q", _dp), ("R", _dp),
                ("S", _dp), ("P", _dp), ("umin", _dp), ("umax", _dp), ("xmin", _dp), ("xmax", _dp),
                ("state_constraint", C.c_int32), ("terminal_mode", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("nx", C.c_int32), ("nu", C.c_int32), ("horizon", C.c_int32), ("nz", C.c_int32), ("mg", C.c_int32),
                ("nt", C.c_int32), ("nt_pad", C.c_int32), ("kernel", C.c_int32), ("device", C.c_int32), ("sm_count", C.c_int32),
                ("rho", C.c_double), ("lambda_min", C.c_double), ("lambda_max", C.c_double)]


class BatchIO(C.Structure):
    _fields_ = [("batch", C.c_int64), ("x0", C.c_void_p), ("xref", C.c_void_p), ("uref", C.c_void_p),
                ("xref_broadcast", C.c_int32), ("uref_broadcast", C.c_int32), ("warm_u", C.c_void_p), ("warm_y", C.c_void_p),
                ("u", C.c_void_p), ("e_u", C.c_void_p), ("x", C.c_void_p), ("e_x", C.c_void_p), ("u0", C.c_void_p),
                ("status", C.c_void_p), ("iters", C.c_void_p), ("prim_res", C.c_void_p), ("dual_res", C.c_void_p),
                ("objective", C.c_void_p), ("y", C.c_void_p), ("inner_iters", C.c_void_p)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("solve_ms", C.c_float), ("recover_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("batch", C.c_int64), ("total_iterations", C.c_int64), ("kernel_launches", C.c_int32), ("chunks", C.c_int32)]


class ClosedLoopIO(C.Structure):
    _fields_ = [("batch", C.c_int64), ("steps", C.c_int32), ("warm_start", C.c_int32), ("x0", C.c_void_p), ("xref", C.c_void_p),
                ("uref", C.c_void_p), ("xref_broadcast", C.c_int32), ("uref_broadcast", C.c_int32), ("x_traj", C.c_void_p),
                ("u_traj", C.c_void_p), ("iters_total", C.c_void_p), ("unsolved_steps", C.c_void_p)]


class NnDesc(C.Structure):
    _fields_ = [("arch", C.c_int32), ("activation", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32), ("n_neurons", C.c_int32),
                ("n_hidden", C.c_int32), ("W_in", _dp), ("W_hidden", _dp), ("b_hidden", _dp), ("W_out", _dp)]


class NmpcDesc(C.Structure):
    _fields_ = [("nn", C.POINTER(NnDesc)), ("horizon", C.c_int32), ("Q", _dp), ("R", _dp), ("S", _dp), ("P", _dp), ("umin", _dp),
                ("umax", _dp), ("xref", _dp), ("uref", _dp), ("terminal_mode", C.c_int32), ("state_constraint", C.c_int32),
                ("xmin", _dp), ("xmax", _dp)]


class NmpcSettings(C.Structure):
    _fields_ = [("qp", Settings), ("sqp_tol", C.c_double), ("ls_armijo", C.c_double), ("ls_noise", C.c_double), ("sqp_max_iter", C.c_int32),
                ("ls_max_halvings", C.c_int32)]


# every symbol include/mpcb200.h declares (tests assert the library exports all of them)
EXPORTED_SYMBOLS = (
    "mpcb_version", "mpcb_device_count", "mpcb_last_error", "mpcb_default_settings", "mpcb_dare", "mpcb_create_linear",
    "mpcb_destroy", "mpcb_get_info", "mpcb_get_timing", "mpcb_get_design", "mpcb_solve_linear_batch",
    "mpcb_solve_linear_batch_device", "mpcb_alloc_pinned", "mpcb_free_pinned", "mpcb_closed_loop_linear_batch",
    "mpcb_create_nn", "mpcb_destroy_nn", "mpcb_nn_rollout_batch", "mpcb_nn_rollout_batch_device", "mpcb_nn_jacobian_batch",
    "mpcb_nn_jacobian_batch_device", "mpcb_default_nmpc_settings", "mpcb_create_nmpc", "mpcb_destroy_nmpc", "mpcb_nmpc_get_design",
    "mpcb_nmpc_get_timing", "mpcb_solve_nmpc_batch", "mpcb_solve_nmpc_batch_device", "mpcb_solve_relinearized_batch",
    "mpcb_solve_relinearized_batch_device", "mpcb_dare_batch", "mpcb_dare_batch_device", "mpcb_closed_loop_nmpc_batch", "mpcb_tune_rho",
)

_lib = None


class MpcbError(RuntimeError):
    pass


def lib():
    """Load libmpcb200.so from the package directory (built in-tree by csrc/build.sh / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise MpcbError(f"{LIB_PATH} is missing: build it with automationlabsmodelpredictivecontrol.jl_b200/csrc/build.sh "
                            "(there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        L.mpcb_last_error.restype = C.c_char_p
        L.mpcb_create_linear.argtypes = [C.POINTER(LinearDesc), C.POINTER(Settings), C.POINTER(C.c_void_p)]
        L.mpcb_destroy.argtypes = [C.c_void_p]
        L.mpcb_destroy.restype = None
        L.mpcb_get_info.argtypes = [C.c_void_p, C.POINTER(Info)]
        L.mpcb_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
        L.mpcb_get_design.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
        L.mpcb_solve_linear_batch.argtypes = [C.c_void_p, C.POINTER(BatchIO)]
        L.mpcb_solve_linear_batch_device.argtypes = [C.c_void_p, C.POINTER(BatchIO), C.c_void_p]
        L.mpcb_closed_loop_linear_batch.argtypes = [C.c_void_p, C.POINTER(ClosedLoopIO)]
        L.mpcb_dare.argtypes = [C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp]
        L.mpcb_tune_rho.argtypes = [C.POINTER(LinearDesc), C.POINTER(Settings), C.POINTER(BatchIO), C.c_int32, C.c_double, C.POINTER(C.c_double), _dp, _dp]
        L.mpcb_alloc_pinned.argtypes = [C.c_size_t]
        L.mpcb_alloc_pinned.restype = C.c_void_p
        L.mpcb_free_pinned.argtypes = [C.c_void_p]
        L.mpcb_free_pinned.restype = None
        L.mpcb_create_nn.argtypes = [C.POINTER(NnDesc), C.c_int32, C.POINTER(C.c_void_p)]
        L.mpcb_destroy_nn.argtypes = [C.c_void_p]
        L.mpcb_destroy_nn.restype = None
        L.mpcb_nn_rollout_batch.argtypes = [C.c_void_p, C.c_int64, C.c_int32, _dp, _dp, _dp]
        L.mpcb_nn_rollout_batch_device.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mpcb_nn_jacobian_batch.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, _dp, _dp]
        L.mpcb_nn_jacobian_batch_device.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6
        L.mpcb_default_nmpc_settings.argtypes = [C.POINTER(NmpcSettings)]
        L.mpcb_default_nmpc_settings.restype = None
        L.mpcb_create_nmpc.argtypes = [C.POINTER(NmpcDesc), C.POINTER(NmpcSettings), C.POINTER(C.c_void_p)]
        L.mpcb_destroy_nmpc.argtypes = [C.c_void_p]
        L.mpcb_destroy_nmpc.restype = None
        L.mpcb_nmpc_get_design.argtypes = [C.c_void_p, C.POINTER(C.c_double), _dp, _dp, _dp]
        L.mpcb_nmpc_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
        L.mpcb_solve_nmpc_batch.argtypes = [C.c_void_p, C.POINTER(BatchIO)]
        L.mpcb_solve_nmpc_batch_device.argtypes = [C.c_void_p, C.POINTER(BatchIO), C.c_void_p]
        L.mpcb_solve_relinearized_batch.argtypes = [C.c_void_p, C.POINTER(BatchIO)]
        L.mpcb_closed_loop_nmpc_batch.argtypes = [C.c_void_p, C.POINTER(ClosedLoopIO)]
        L.mpcb_solve_relinearized_batch_device.argtypes = [C.c_void_p, C.POINTER(BatchIO), C.c_void_p]
        L.mpcb_dare_batch.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp, C.POINTER(C.c_int32)]
        L.mpcb_dare_batch_device.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc, what):
    if rc != MPCB_OK:
        msg = lib().mpcb_last_error()
        raise MpcbError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def default_nmpc_settings(**kw) -> NmpcSettings:
    """kw: sqp_tol, ls_armijo, ls_noise, sqp_max_iter, ls_max_halvings, or any inner-QP setting of `Settings`."""
    s = NmpcSettings()
    lib().mpcb_default_nmpc_settings(C.byref(s))
    for k, v in kw.items():
        if hasattr(s, k) and k != "qp": setattr(s, k, v)
        elif k == "devices":
            ids = [int(d) for d in v]
            if not 1 <= len(ids) <= 8: raise ValueError("devices: 1..8 device ordinals")
            s.qp.n_devices = len(ids); s.qp.device = ids[0]
            for i, d in enumerate(ids): s.qp.device_ids[i] = d
        elif hasattr(s.qp, k): setattr(s.qp, k, v)
        else: raise TypeError(f"unknown NMPC setting {k!r}")
    return s


def default_settings(**kw) -> Settings:
    s = Settings()
    lib().mpcb_default_settings(C.byref(s))
    for k, v in kw.items():
        if k == "devices":          # devices=[0, 1, ...]: one handle driving several GPUs from this process
            ids = [int(d) for d in v]
            if not 1 <= len(ids) <= 8: raise ValueError("devices: 1..8 device ordinals")
            s.n_devices = len(ids); s.device = ids[0]
            for i, d in enumerate(ids): s.device_ids[i] = d
            continue
        if not hasattr(s, k):
            raise TypeError(f"unknown solver setting {k!r}")
        setattr(s, k, v)
    return s
