"""B200NonlinearModeler: what `tuning.modeler` holds for `mpc_programming_type = "non_linear"` with `mpc_solver = "b200"`
-- the NL modeler of the reference (fnn.jl:63-189, resnet.jl:62-188) + Ipopt, replaced by the batched SQP kernel."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_NMPC_SETTING_KEYS = {"mpc_b200_sqp_tol": "sqp_tol", "mpc_b200_sqp_max_iter": "sqp_max_iter", "mpc_b200_ls_armijo": "ls_armijo", "mpc_b200_ls_noise": "ls_noise",
                      "mpc_b200_ls_max_halvings": "ls_max_halvings", "mpc_b200_eps_abs": "eps_abs", "mpc_b200_eps_rel": "eps_rel",
                      "mpc_b200_rho": "rho", "mpc_b200_max_iter": "max_iter", "mpc_b200_check_every": "check_every", "mpc_b200_device": "device",
                      "mpc_b200_alpha": "alpha", "mpc_b200_sigma": "sigma", "mpc_b200_devices": "devices"}


class B200NonlinearModeler:
    def __init__(self, nn, Q, R, S, P, umin, umax, xmin, xmax, horizon, xref, uref, state_constraint=False, terminal="none", kws=None):
        if terminal not in ("none", "equality", "contractive"):
            raise _lib.MpcbError(f"mpc_terminal_ingredient={terminal!r} is not supported by mpc_solver='b200' (only 'none', 'equality' and 'contractive')")
        kws = kws or {}
        self.nn = nn
        self.nx, self.nu, self.horizon = nn.nx, nn.nu, int(horizon)
        self.settings = _lib.default_nmpc_settings(**{v: kws[k] for k, v in _NMPC_SETTING_KEYS.items() if k in kws})
        nd, keep_nn = nn.desc()
        f = lambda a: None if a is None else np.asfortranarray(np.asarray(a, np.float64))
        keep = [f(Q), f(R), f(S), f(P), f(umin), f(umax), f(xref), f(uref)]
        p = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))
        self.terminal = terminal
        keep_x = [f(xmin), f(xmax)]
        d = _lib.NmpcDesc(C.pointer(nd), self.horizon, *[p(a) for a in keep],
                          {"equality": _lib.TERMINAL_EQUALITY, "contractive": _lib.TERMINAL_CONTRACTIVE}.get(terminal, _lib.TERMINAL_NONE), 1 if state_constraint else 0,
                          p(keep_x[0]), p(keep_x[1]))
        self._h = C.c_void_p()
        _lib.check(_lib.lib().mpcb_create_nmpc(C.byref(d), C.byref(self.settings), C.byref(self._h)), "mpcb_create_nmpc")
        del keep_nn
        self.nz = self.nu * self.horizon
        self.state_constraint = bool(state_constraint)
        # duals: input box rows [+ state-box rows] [+ terminal rows]
        self.ny = self.nz + (self.nx * self.horizon if state_constraint else 0) + (self.nx if terminal != "none" else 0)
        self.x0 = self.xref = self.uref = None
        self.warm = None

    def design(self):
        rho = C.c_double(); A = np.zeros((self.nx, self.nx), order="F"); B = np.zeros((self.nx, self.nu), order="F"); P = np.zeros((self.nx, self.nx), order="F")
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        _lib.check(_lib.lib().mpcb_nmpc_get_design(self._h, C.byref(rho), p(A), p(B), p(P)), "mpcb_nmpc_get_design")
        return {"rho": rho.value, "A": np.ascontiguousarray(A), "B": np.ascontiguousarray(B), "P": np.ascontiguousarray(P)}

    def timing(self):
        t = _lib.Timing()
        _lib.check(_lib.lib().mpcb_nmpc_get_timing(self._h, C.byref(t)), "mpcb_nmpc_get_timing")
        return {k: getattr(t, k) for k, _ in t._fields_}

    def solve_batch(self, x0, xref, uref, want=("u", "e_u", "x", "e_x", "u0", "objective"), warm=None, out=None, method="non_linear"):
        """Host-array entry.  method = "non_linear": the SQP solve of the NL modeler (mpcb_solve_nmpc_batch);
        method = "linear": the reference's linear method on the black-box model (design_mpc.jl:319-327), re-designed per
        problem at that problem's reference on the device (mpcb_solve_relinearized_batch).
        warm = (u_init or None, y_init or None)."""
        if method not in ("non_linear", "linear"): raise ValueError("method must be 'non_linear' or 'linear'")
        x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, np.float64)))
        Bn = x0.shape[0]
        if x0.shape[1] != self.nx: raise ValueError("x0 must be (batch, nx)")
        xref = np.ascontiguousarray(np.asarray(xref, np.float64)); uref = np.ascontiguousarray(np.asarray(uref, np.float64))
        xb = xref.ndim == 1 or xref.shape[0] == 1 and Bn != 1
        ub = uref.ndim == 1 or uref.shape[0] == 1 and Bn != 1
        if xref.size != (self.nx if xb else self.nx * Bn) or uref.size != (self.nu if ub else self.nu * Bn):
            raise ValueError("reference shapes do not match the batch")
        H = self.horizon
        shapes = {"u": (Bn, H, self.nu), "e_u": (Bn, H, self.nu), "x": (Bn, H + 1, self.nx), "e_x": (Bn, H + 1, self.nx), "u0": (Bn, self.nu),
                  "objective": (Bn,), "prim_res": (Bn,), "dual_res": (Bn,), "y": (Bn, self.ny)}
        res = {} if out is None else out
        for k in tuple(want) + ("prim_res", "dual_res"):
            if k not in res: res[k] = np.empty(shapes[k], np.float64)
        for k in ("status", "iters", "inner_iters"):
            if k not in res: res[k] = np.empty(Bn, np.int32)
        io = _lib.BatchIO()
        io.batch = Bn; io.x0 = x0.ctypes.data; io.xref = xref.ctypes.data; io.uref = uref.ctypes.data
        io.xref_broadcast = int(xb); io.uref_broadcast = int(ub)
        keep = []
        if warm is not None:
            if warm[0] is not None:
                wu = np.ascontiguousarray(warm[0], np.float64); keep.append(wu)
                if wu.size != Bn * self.nz: raise ValueError("warm start shapes")
                io.warm_u = wu.ctypes.data
            if warm[1] is not None:
                wy = np.ascontiguousarray(warm[1], np.float64); keep.append(wy)
                if wy.size != Bn * self.ny: raise ValueError("warm start shapes")
                io.warm_y = wy.ctypes.data
        for k in ("u", "e_u", "x", "e_x", "u0", "objective", "prim_res", "dual_res", "y", "status", "iters", "inner_iters"):
            if k in res: setattr(io, k, res[k].ctypes.data)
        if method == "linear": _lib.check(_lib.lib().mpcb_solve_relinearized_batch(self._h, C.byref(io)), "mpcb_solve_relinearized_batch")
        else: _lib.check(_lib.lib().mpcb_solve_nmpc_batch(self._h, C.byref(io)), "mpcb_solve_nmpc_batch")
        return res

    def solve_batch_device(self, io: _lib.BatchIO, stream=None, method="non_linear"):
        if method == "linear":
            _lib.check(_lib.lib().mpcb_solve_relinearized_batch_device(self._h, C.byref(io), C.c_void_p(stream or 0)), "mpcb_solve_relinearized_batch_device")
        else:
            _lib.check(_lib.lib().mpcb_solve_nmpc_batch_device(self._h, C.byref(io), C.c_void_p(stream or 0)), "mpcb_solve_nmpc_batch_device")

    def closed_loop(self, x0, xref, uref, steps, warm_start=True):
        """GPU-resident closed loop on the network as the plant (mpcb_closed_loop_nmpc_batch): x0 (B, nx) -> x_traj (B, steps+1, nx),
        u_traj (B, steps, nu), iters_total (B,) inner ADMM iterations, unsolved_steps (B,)."""
        x0 = np.ascontiguousarray(np.atleast_2d(np.asarray(x0, np.float64))); Bn = x0.shape[0]
        xref = np.ascontiguousarray(np.asarray(xref, np.float64)); uref = np.ascontiguousarray(np.asarray(uref, np.float64))
        xb = xref.ndim == 1 or xref.shape[0] == 1 and Bn != 1
        ub = uref.ndim == 1 or uref.shape[0] == 1 and Bn != 1
        if x0.shape[1] != self.nx or xref.size != (self.nx if xb else self.nx * Bn) or uref.size != (self.nu if ub else self.nu * Bn):
            raise ValueError("closed_loop: shapes")
        out = {"x_traj": np.empty((Bn, steps + 1, self.nx)), "u_traj": np.empty((Bn, steps, self.nu)), "iters_total": np.empty(Bn, np.int32),
               "unsolved_steps": np.empty(Bn, np.int32)}
        io = _lib.ClosedLoopIO(Bn, int(steps), int(bool(warm_start)), x0.ctypes.data, xref.ctypes.data, uref.ctypes.data, int(xb), int(ub),
                               out["x_traj"].ctypes.data, out["u_traj"].ctypes.data, out["iters_total"].ctypes.data, out["unsolved_steps"].ctypes.data)
        _lib.check(_lib.lib().mpcb_closed_loop_nmpc_batch(self._h, C.byref(io)), "mpcb_closed_loop_nmpc_batch")
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().mpcb_destroy_nmpc(self._h); self._h = None

    def __del__(self):
        try: self.close()
        except Exception: pass
