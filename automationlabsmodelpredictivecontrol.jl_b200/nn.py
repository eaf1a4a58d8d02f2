"""Neural dynamics models of the nonlinear path: what `system.f` holds for a ConstrainedBlackBoxControlDiscreteSystem.

The reference stores a Flux `Chain` there and its NL modelers read `Flux.params(system.f)` positionally
(/root/reference/src/sub/model_modeler_implementation/fnn/mpc_modeler_implementation_fnn.jl:88-107): params[1] = W_in (no
bias), then (W_j, b_j) pairs for the hidden layers, params[end] = W_out (no bias); the activation is read off the first
hidden layer (src/sub/design_mpc.jl:472-496).  `Fnn` / `ResNet` carry exactly that, plus a handle to the copy of the
weights resident on the GPU (mpcb_create_nn).  There is no CPU evaluation here: `__call__`, `rollout` and `jacobian` run
the CUDA kernels through the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class _Chain:
    arch = None      # "fnn" | "resnet" | "polynet"

    def __init__(self, W_in, hidden, W_out, activation="relu", device=0):
        """hidden: sequence of (W_j, b_j); layouts as Flux stores them (out x in)."""
        if activation not in _lib.ACTIVATION_IDS:
            raise ValueError(f"unsupported activation {activation!r} (supported: {sorted(_lib.ACTIVATION_IDS)})")
        self.activation = activation
        self.W_in = _f(W_in); self.W_out = _f(W_out)
        self.W_h = [_f(W) for W, _ in hidden]; self.b_h = [np.ascontiguousarray(b, np.float64) for _, b in hidden]
        self.n_neurons, nin = self.W_in.shape
        self.nx = self.W_out.shape[0]; self.nu = nin - self.nx
        n, nh = self.n_neurons, len(self.W_h)
        dense = self.arch == "densenet"          # layer j sees the concatenation of all earlier blocks
        ok_h = all(W.shape == (n, (l + 1) * n if dense else n) for l, W in enumerate(self.W_h))
        if self.nu <= 0 or self.W_out.shape[1] != ((nh + 1) * n if dense else n) or not ok_h or any(b.shape != (n,) for b in self.b_h):
            raise ValueError("inconsistent layer shapes")
        self.device = device
        self._h = None

    # -- Flux.params(system.f) order -----------------------------------------------------------------------------
    def params(self):
        out = [self.W_in]
        for W, b in zip(self.W_h, self.b_h): out += [W, b]
        return out + [self.W_out]

    # -- C ABI -------------------------------------------------------------------------------------------------------
    def desc(self):
        """(NnDesc, keepalive) -- the arrays must outlive the call that consumes the struct."""
        nh = len(self.W_h)
        Wh = np.concatenate([W.ravel(order="F") for W in self.W_h]) if nh else np.zeros(1)
        bh = np.concatenate(self.b_h) if nh else np.zeros(1)
        keep = [self.W_in, Wh, bh, self.W_out]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        d = _lib.NnDesc({"fnn": _lib.NN_FNN, "resnet": _lib.NN_RESNET, "polynet": _lib.NN_POLYNET, "densenet": _lib.NN_DENSENET}[self.arch], _lib.ACTIVATION_IDS[self.activation], self.nx, self.nu,
                        self.n_neurons, nh, p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]))
        return d, keep

    def handle(self):
        if self._h is None:
            d, keep = self.desc()
            h = C.c_void_p()
            _lib.check(_lib.lib().mpcb_create_nn(C.byref(d), self.device, C.byref(h)), "mpcb_create_nn")
            self._h = h
        return self._h

    def rollout(self, x0, u):
        """x0 (B, nx), u (B, H, nu) -> x (B, H+1, nx) on the GPU (mpcb_nn_rollout_batch)."""
        x0 = np.ascontiguousarray(np.atleast_2d(x0), np.float64); u = np.ascontiguousarray(u, np.float64)
        if u.ndim == 2: u = u[None]
        Bn, H = u.shape[0], u.shape[1]
        if x0.shape != (Bn, self.nx) or u.shape[2] != self.nu: raise ValueError("rollout: shapes")
        x = np.empty((Bn, H + 1, self.nx))
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        _lib.check(_lib.lib().mpcb_nn_rollout_batch(self.handle(), Bn, H, p(x0), p(u), p(x)), "mpcb_nn_rollout_batch")
        return x

    def jacobian(self, x, u):
        """x (B, nx), u (B, nu) -> f (B, nx), A (B, nx, nx), Bm (B, nx, nu) on the GPU (mpcb_nn_jacobian_batch)."""
        x = np.ascontiguousarray(np.atleast_2d(x), np.float64); u = np.ascontiguousarray(np.atleast_2d(u), np.float64)
        Bn = x.shape[0]
        if x.shape != (Bn, self.nx) or u.shape != (Bn, self.nu): raise ValueError("jacobian: shapes")
        f = np.empty((Bn, self.nx)); A = np.empty((Bn, self.nx, self.nx)); Bm = np.empty((Bn, self.nu, self.nx))
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        _lib.check(_lib.lib().mpcb_nn_jacobian_batch(self.handle(), Bn, p(x), p(u), p(f), p(A), p(Bm)), "mpcb_nn_jacobian_batch")
        return f, A.transpose(0, 2, 1).copy(), Bm.transpose(0, 2, 1).copy()      # column-major per problem -> [row, col]

    def __call__(self, xu):
        """system.f([x; u]) for one point, as the reference calls the Chain."""
        xu = np.asarray(xu, np.float64).ravel()
        return self.rollout(xu[None, :self.nx], xu[None, None, self.nx:])[0, 1]

    def close(self):
        if self._h is not None:
            _lib.lib().mpcb_destroy_nn(self._h); self._h = None

    def __del__(self):
        try: self.close()
        except Exception: pass


class Fnn(_Chain):
    """AutomationLabsSystems.Fnn: y_j = act(W_j y_{j-1} + b_j)  (fnn.jl:133-141)."""
    arch = "fnn"


class Icnn(Fnn):
    """AutomationLabsSystems.Icnn: the reference's NL / linear modelers for it are the Fnn ones up to the type tag
    (icnn/mpc_modeler_implementation_icnn.jl is fnn.jl with `Fnn` -> `Icnn`)."""


class Rbf(Fnn):
    """AutomationLabsSystems.Rbf: rbf/mpc_modeler_implementation_rbf.jl:23-58 (linear) and :61-186 (NL) are the Fnn modelers
    (no MILP variant)."""


class ResNet(_Chain):
    """AutomationLabsSystems.ResNet: y_j = y_{j-1} + act(W_j y_{j-1} + b_j)  (resnet.jl:131-140)."""
    arch = "resnet"


class PolyNet(_Chain):
    """AutomationLabsSystems.PolyNet: br = act(W_j y + b_j); y_j = y_{j-1} + br + act(W_j br + b_j)  (polynet.jl:132-149)."""
    arch = "polynet"


class DenseNet(_Chain):
    """AutomationLabsSystems.DenseNet: y_j = [act(W_j y_{j-1} + b_{j-1}); y_{j-1}], W_j of size n x ((j-1) n), W_out nx x ((n_hidden+1) n)
    (densenet.jl:128-162)."""
    arch = "densenet"


def linearize(nn: _Chain, x, u):
    """AutomationLabsSystems.proceed_system_linearization (fnn.jl:42, design_mpc.jl:319-323): Jacobians at (x, u)."""
    _, A, B = nn.jacobian(np.asarray(x, float)[None], np.asarray(u, float)[None])
    return A[0], B[0]
