"""Design orchestration, mirroring /root/reference/src/sub/design_mpc.jl (line numbers cited per function).

The reference builds a JuMP model here; with mpc_solver="b200" the same inputs (system, horizon, references, kws)
produce a B200Modeler (condensed QP + cached KKT operator on the GPU) and the SAME controller structs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .modeler import B200Modeler
from .solver_selection import _IMPLEMENTATION_SOLVER_LIST, require_b200, resolve_solver
from .systems import ConstrainedBlackBoxControlDiscreteSystem, ConstrainedLinearControlDiscreteSystem
from .types import (IMPLEMENTATION_PROGRAMMING_LIST, LinearProgramming, ModelPredictiveControlController,
                    ModelPredictiveControlResults, ModelPredictiveControlTuning, NonLinearProgramming, ReferencesStateInput,
                    TerminalIngredient, WeightsCoefficient)

# main_mpc.jl:87-94
_DEFAULT_PARAMETERS_MODEL_PREDICTIVE_CONTROL = dict(mpc_solver="auto", mpc_terminal_ingredient="none", mpc_Q=100.0, mpc_R=0.1,
                                                    mpc_S=0.0, mpc_max_time=30.0)

# solver settings that may be forwarded through the reference's kwargs mechanism (unknown keys are silently ignored by
# the reference, design_mpc.jl:63-64; these are the new, recognised ones)
_B200_SETTING_KEYS = {"mpc_b200_eps_abs": "eps_abs", "mpc_b200_eps_rel": "eps_rel", "mpc_b200_rho": "rho", "mpc_b200_max_iter": "max_iter",
                      "mpc_b200_check_every": "check_every", "mpc_b200_device": "device", "mpc_b200_kernel": "kernel",
                      "mpc_b200_alpha": "alpha", "mpc_b200_sigma": "sigma", "mpc_b200_eps_prim_inf": "eps_prim_inf",
                      "mpc_b200_ladder_iter": "ladder_iter", "mpc_b200_ladder_kappa": "ladder_kappa", "mpc_b200_devices": "devices",
                      "mpc_b200_cold_init": "cold_init"}


def _settings_from_kws(kws) -> _lib.Settings:
    return _lib.default_settings(**{v: kws[k] for k, v in _B200_SETTING_KEYS.items() if k in kws})


def dare(A, B, Q, R):
    """P = are(Discrete, A, B, Q, R)  (design_mpc.jl:327) through libmpcb200's doubling solver."""
    A = np.asfortranarray(A, dtype=np.float64); B = np.asfortranarray(B, dtype=np.float64)
    Q = np.asfortranarray(Q, dtype=np.float64); R = np.asfortranarray(R, dtype=np.float64)
    nx, nu = B.shape
    P = np.zeros((nx, nx), order="F")
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.check(_lib.lib().mpcb_dare(nx, nu, p(A), p(B), p(Q), p(R), p(P)), "mpcb_dare")
    return np.ascontiguousarray(P)


def dare_batch(A, B, Q, R, device=0):
    """P_i = are(Discrete, A_i, B_i, Q, R) for many systems at once on the GPU (mpcb_dare_batch): A (n, nx, nx), B (n, nx, nu).
    Returns (P (n, nx, nx), steps (n,)): steps > 0 doubling steps taken, -1 where the equation has no stabilising solution."""
    A = np.asarray(A, np.float64); B = np.asarray(B, np.float64)
    n, nx, nu = B.shape
    Af = np.ascontiguousarray(A.transpose(0, 2, 1)); Bf = np.ascontiguousarray(B.transpose(0, 2, 1))     # column-major per system
    Q = np.asfortranarray(Q, dtype=np.float64); R = np.asfortranarray(R, dtype=np.float64)
    P = np.empty((n, nx, nx)); st = np.empty(n, np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.check(_lib.lib().mpcb_dare_batch(device, n, nx, nu, p(Af), p(Bf), p(Q), p(R), p(P), st.ctypes.data_as(C.POINTER(C.c_int32))), "mpcb_dare_batch")
    return P.transpose(0, 2, 1).copy(), st


def _create_weights_coefficients(system, kws) -> WeightsCoefficient:
    """design_mpc.jl:235-283: scalar * identity for Q, R, S."""
    d = _DEFAULT_PARAMETERS_MODEL_PREDICTIVE_CONTROL
    Q = kws.get("mpc_Q", d["mpc_Q"]); R = kws.get("mpc_R", d["mpc_R"]); S = kws.get("mpc_S", d["mpc_S"])
    nx, nu = system.statedim, system.inputdim
    return WeightsCoefficient(Q * np.eye(nx), R * np.eye(nu), S * np.eye(nu))


def _bounds(system):
    """linear.jl:34-38: first/last vertex of the hyperrectangles == low/high corners."""
    return system.U.low, system.U.high, system.X.low, system.X.high


def _memory_allocation_initialization_results_mpc(system, horizon):
    """design_mpc.jl:499-529 (Julia leaves the buffers undef; zeros here)."""
    nx, nu = system.statedim, system.inputdim
    return np.zeros(nx), ModelPredictiveControlResults(np.zeros((nx, horizon + 1)), np.zeros((nx, horizon + 1)),
                                                       np.zeros((nu, horizon)), np.zeros((nu, horizon)))


def _model_predictive_control_design(system, horizon: int, sample_time: int, references: ReferencesStateInput, **kws):
    """design_mpc.jl:54-129 (linear discrete) and :143-225 (black box)."""
    d = _DEFAULT_PARAMETERS_MODEL_PREDICTIVE_CONTROL
    linear_system = isinstance(system, ConstrainedLinearControlDiscreteSystem)
    if not linear_system and not isinstance(system, ConstrainedBlackBoxControlDiscreteSystem):
        raise TypeError(f"unsupported system type {type(system).__name__}")
    programming = kws.get("mpc_programming_type", "linear" if linear_system else "non_linear")   # :67 / :159
    method = IMPLEMENTATION_PROGRAMMING_LIST[programming]
    solver = resolve_solver(method, _IMPLEMENTATION_SOLVER_LIST[kws.get("mpc_solver", d["mpc_solver"])])
    terminal = kws.get("mpc_terminal_ingredient", d["mpc_terminal_ingredient"])
    max_time = kws.get("mpc_max_time", d["mpc_max_time"])
    require_b200(solver)
    weights = _create_weights_coefficients(system, kws)
    umin, umax, xmin, xmax = _bounds(system)
    state_constraint = "mpc_state_constraint" in kws          # presence-only flag, linear.jl:62

    if linear_system:
        if not isinstance(method, LinearProgramming):
            raise TypeError("a linear system only has the LinearProgramming modeler (linear.jl:20-27)")
        A, B = system.A, system.B
        A_term, B_term = A, B
        nn = None
    else:
        from . import nn as nnmod
        nn = system.f
        # linearise at the FIRST reference column for the dynamics (fnn.jl:38-46) and at the LAST for P
        # (design_mpc.jl:312-323); identical for the constant references proceed_controller builds
        A, B = nnmod.linearize(nn, references.x[:, 0], references.u[:, 0])
        A_term, B_term = nnmod.linearize(nn, references.x[:, -1], references.u[:, -1])

    P = dare(A_term, B_term, weights.Q, weights.R)            # design_mpc.jl:327: always, even for Xf = "none"
    settings = _settings_from_kws(kws)
    if isinstance(method, LinearProgramming):
        modeler = B200Modeler(A, B, weights.Q, weights.R, weights.S, P, umin, umax, xmin, xmax, horizon,
                              state_constraint=state_constraint, terminal=terminal, settings=settings, rho_tune=kws.get("mpc_b200_rho_tune"))
    elif isinstance(method, NonLinearProgramming):
        from .nmpc import B200NonlinearModeler
        modeler = B200NonlinearModeler(nn, weights.Q, weights.R, weights.S, P, umin, umax, xmin, xmax, horizon,
                                       references.x[:, -1], references.u[:, -1], state_constraint=state_constraint,
                                       terminal=terminal, kws=kws)
    else:
        raise NotImplementedError("mixed-integer / fuzzy programming is outside the B200 path (SURVEY.md section 2)")
    tuning = ModelPredictiveControlTuning(modeler, references, horizon, weights, TerminalIngredient(terminal, P),
                                          float(sample_time), int(max_time))
    initialization, results = _memory_allocation_initialization_results_mpc(system, horizon)
    return ModelPredictiveControlController(system, tuning, initialization, results)
