"""B200Modeler: what `tuning.modeler` holds when `mpc_solver = "b200"` -- the condensed design plus an opaque
libmpcb200 handle, in place of the JuMP model the reference stores there (src/types/types.jl:115)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _colmajor(a, shape=None):
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return np.asfortranarray(a)


class B200Modeler:
    def __init__(self, A, B, Q, R, S, P, umin, umax, xmin, xmax, horizon, state_constraint=False, terminal="none",
                 settings: _lib.Settings | None = None, rho_tune=None):
        L = _lib.lib()
        self.nx, self.nu = np.asarray(B).shape
        self.horizon = int(horizon)
        if terminal not in ("none", "equality", "contractive"):
            # "neighborhood" is a @warn stub in the reference (design_mpc.jl:342-345)
            raise _lib.MpcbError(f"mpc_terminal_ingredient={terminal!r} is not supported by mpc_solver='b200' "
                                 "(only 'none', 'equality' and 'contractive')")
        keep = [_colmajor(A), _colmajor(B), _colmajor(Q), _colmajor(R), _colmajor(S),
                None if P is None else _colmajor(P), _colmajor(umin), _colmajor(umax),
                None if xmin is None else _colmajor(xmin), None if xmax is None else _colmajor(xmax)]
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
        d = _lib.LinearDesc(self.nx, self.nu, self.horizon, *[ptr(a) for a in keep], 1 if state_constraint else 0,
                            {"none": _lib.TERMINAL_NONE, "equality": _lib.TERMINAL_EQUALITY, "contractive": _lib.TERMINAL_CONTRACTIVE}[terminal])
        self.settings = settings if settings is not None else _lib.default_settings()
        self.rho_tuning = None
        if rho_tune is not None:
            # kw `mpc_b200_rho_tune = (x0_sample, xref, uref)`: pick the step size on a sample of the workload (mpcb_tune_rho) -- the batch-wide
            # counterpart of OSQP's per-problem adaptive rho; a design-time step
            x0s = np.ascontiguousarray(np.atleast_2d(np.asarray(rho_tune[0], np.float64)))
            xrs = np.ascontiguousarray(np.asarray(rho_tune[1], np.float64)); urs = np.ascontiguousarray(np.asarray(rho_tune[2], np.float64))
            io = _lib.BatchIO(); io.batch = x0s.shape[0]; io.x0 = x0s.ctypes.data; io.xref = xrs.ctypes.data; io.uref = urs.ctypes.data
            io.xref_broadcast = int(xrs.ndim == 1); io.uref_broadcast = int(urs.ndim == 1)
            ncand = int(rho_tune[3]) if len(rho_tune) > 3 else 7
            best = C.c_double(); cr = np.zeros(ncand); ci = np.zeros(ncand)
            _lib.check(L.mpcb_tune_rho(C.byref(d), C.byref(self.settings), C.byref(io), ncand, 2.0, C.byref(best), cr.ctypes.data_as(C.POINTER(C.c_double)),
                                       ci.ctypes.data_as(C.POINTER(C.c_double))), "mpcb_tune_rho")
            self.settings.rho = best.value
            self.rho_tuning = {"rho": best.value, "candidates": cr.tolist(), "mean_iters": ci.tolist()}
        self._h = C.c_void_p()
        _lib.check(L.mpcb_create_linear(C.byref(d), C.byref(self.settings), C.byref(self._h)), "mpcb_create_linear")
        self.info = _lib.Info()
        _lib.check(L.mpcb_get_info(self._h, C.byref(self.info)), "mpcb_get_info")
        self.x0 = None          # (B, nx) set by update_initialization
        self.xref = None
        self.uref = None
        self.warm = None        # (u, y) of the previous solve, for closed-loop warm starts

    # -- introspection -------------------------------------------------------------------------------------------
    def design(self):
        i = self.info; npar = 2 * i.nx + i.nu
        Pc = np.zeros((i.nz, i.nz), order="F"); Lq = np.zeros((i.nz, npar), order="F")
        G = np.zeros((i.mg, i.nz), order="F"); Lb = np.zeros((i.mg, npar), order="F"); T = np.zeros((i.nt, i.nt), order="F")
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        _lib.check(_lib.lib().mpcb_get_design(self._h, p(Pc), p(Lq), p(G), p(Lb), p(T)), "mpcb_get_design")
        return {"Pc": Pc, "Lq": Lq, "G": G, "Lb": Lb, "T": T}

    def timing(self):
        t = _lib.Timing()
        _lib.check(_lib.lib().mpcb_get_timing(self._h, C.byref(t)), "mpcb_get_timing")
        return {k: getattr(t, k) for k, _ in t._fields_}

    # -- the hot path --------------------------------------------------------------------------------------------
    def solve_batch(self, x0, xref, uref, want=("u", "e_u", "x", "e_x", "u0", "objective"), warm=None, out=None):
        """Host-array entry: x0 (B,nx); xref (nx,) or (B,nx); uref (nu,) or (B,nu).  Returns dict of numpy arrays in
        the reference's per-problem layout (row-major (B,H,nu) == column-major nu x H x B)."""
        i = self.info
        x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, np.float64)))
        Bn = x0.shape[0]
        if x0.shape[1] != i.nx: raise ValueError("x0 must be (batch, nx)")
        xref = np.ascontiguousarray(np.asarray(xref, np.float64)); uref = np.ascontiguousarray(np.asarray(uref, np.float64))
        xb = xref.ndim == 1 or xref.shape[0] == 1 and Bn != 1
        ub = uref.ndim == 1 or uref.shape[0] == 1 and Bn != 1
        if xref.size != (i.nx if xb else i.nx * Bn) or uref.size != (i.nu if ub else i.nu * Bn):
            raise ValueError("reference shapes do not match the batch")
        shapes = {"u": (Bn, i.horizon, i.nu), "e_u": (Bn, i.horizon, i.nu), "x": (Bn, i.horizon + 1, i.nx),
                  "e_x": (Bn, i.horizon + 1, i.nx), "u0": (Bn, i.nu), "objective": (Bn,), "prim_res": (Bn,), "dual_res": (Bn,),
                  "y": (Bn, i.nt)}
        res = {} if out is None else out
        for k in tuple(want) + ("prim_res", "dual_res"):
            if k not in res: res[k] = np.empty(shapes[k], np.float64)
        for k in ("status", "iters"):
            if k not in res: res[k] = np.empty(Bn, np.int32)
        io = _lib.BatchIO()
        io.batch = Bn
        io.x0 = x0.ctypes.data; io.xref = xref.ctypes.data; io.uref = uref.ctypes.data
        io.xref_broadcast = int(xb); io.uref_broadcast = int(ub)
        if warm is not None:
            wu = np.ascontiguousarray(warm[0], np.float64); wy = np.ascontiguousarray(warm[1], np.float64)
            if wu.size != Bn * i.nz or wy.size != Bn * i.nt: raise ValueError("warm start shapes")
            io.warm_u = wu.ctypes.data; io.warm_y = wy.ctypes.data
        for k in ("u", "e_u", "x", "e_x", "u0", "objective", "prim_res", "dual_res", "y", "status", "iters"):
            if k in res: setattr(io, k, res[k].ctypes.data)
        _lib.check(_lib.lib().mpcb_solve_linear_batch(self._h, C.byref(io)), "mpcb_solve_linear_batch")
        return res

    def solve_batch_device(self, io: _lib.BatchIO, stream=None):
        """Device-pointer entry (torch tensors' data_ptr()); asynchronous on `stream`."""
        _lib.check(_lib.lib().mpcb_solve_linear_batch_device(self._h, C.byref(io), C.c_void_p(stream or 0)),
                   "mpcb_solve_linear_batch_device")

    def closed_loop(self, x0, xref, uref, steps, warm_start=True):
        """GPU-resident closed loop (mpcb_closed_loop_linear_batch): x0 (B, nx) -> x_traj (B, steps+1, nx), u_traj (B, steps, nu),
        iters_total (B,), unsolved_steps (B,)."""
        i = self.info
        x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, np.float64))); Bn = x0.shape[0]
        xref = np.ascontiguousarray(np.asarray(xref, np.float64)); uref = np.ascontiguousarray(np.asarray(uref, np.float64))
        xb = xref.ndim == 1 or xref.shape[0] == 1 and Bn != 1
        ub = uref.ndim == 1 or uref.shape[0] == 1 and Bn != 1
        if x0.shape[1] != i.nx or xref.size != (i.nx if xb else i.nx * Bn) or uref.size != (i.nu if ub else i.nu * Bn):
            raise ValueError("closed_loop: shapes")
        out = {"x_traj": np.empty((Bn, steps + 1, i.nx)), "u_traj": np.empty((Bn, steps, i.nu)), "iters_total": np.empty(Bn, np.int32),
               "unsolved_steps": np.empty(Bn, np.int32)}
        io = _lib.ClosedLoopIO(Bn, int(steps), int(bool(warm_start)), x0.ctypes.data, xref.ctypes.data, uref.ctypes.data, int(xb), int(ub),
                               out["x_traj"].ctypes.data, out["u_traj"].ctypes.data, out["iters_total"].ctypes.data, out["unsolved_steps"].ctypes.data)
        _lib.check(_lib.lib().mpcb_closed_loop_linear_batch(self._h, C.byref(io)), "mpcb_closed_loop_linear_batch")
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().mpcb_destroy(self._h); self._h = None

    def __del__(self):
        try: self.close()
        except Exception: pass
