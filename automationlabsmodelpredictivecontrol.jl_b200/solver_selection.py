"""Solver table, mirroring /root/reference/src/sub/solver_selection.jl.

The reference maps (method, solver tag) -> an empty `JuMP.Model(optimizer)` (solver_selection.jl:18-114).  Here only
the new tag `b200` constructs anything; the CPU tags are kept in the table so the legality matrix and the `auto`
resolution can be asserted exactly as the reference's tests do, but selecting them raises: this package is the
B200 path only and has no CPU solver behind it."""
from __future__ import annotations

from .types import (AbstractSolvers, LinearProgramming, MixedIntegerLinearProgramming, NonLinearProgramming,
                    auto_solver_def, b200_solver_def, ipopt_solver_def, osqp_solver_def, scip_solver_def)

_IMPLEMENTATION_SOLVER_LIST = {           # solver_selection.jl:9-14 (+ b200)
    "osqp": osqp_solver_def(),
    "scip": scip_solver_def(),
    "ipopt": ipopt_solver_def(),
    "auto": auto_solver_def(),
    "b200": b200_solver_def(),
}

# legality matrix of _JuMP_model_definition (solver_selection.jl:18-53) with b200 added to linear + non-linear
_LEGAL = {
    LinearProgramming: (osqp_solver_def, scip_solver_def, ipopt_solver_def, b200_solver_def),
    NonLinearProgramming: (ipopt_solver_def, scip_solver_def, b200_solver_def),
    MixedIntegerLinearProgramming: (scip_solver_def,),
}
# `auto` resolution (solver_selection.jl:56-87): linear -> scip, non-linear -> ipopt, MILP -> scip
_AUTO = {LinearProgramming: scip_solver_def, NonLinearProgramming: ipopt_solver_def, MixedIntegerLinearProgramming: scip_solver_def}


def resolve_solver(method, solver: AbstractSolvers) -> AbstractSolvers:
    if isinstance(solver, auto_solver_def):
        solver = _AUTO[type(method)]()
    if type(method) not in _LEGAL or not isinstance(solver, _LEGAL[type(method)]):
        # the reference raises a MethodError here (no matching _JuMP_model_definition method)
        raise TypeError(f"no method matching _JuMP_model_definition(::{type(method).__name__}, ::{type(solver).__name__})")
    return solver


def solver_name(solver: AbstractSolvers) -> str:
    return {osqp_solver_def: "OSQP", scip_solver_def: "SCIP", ipopt_solver_def: "Ipopt", b200_solver_def: "B200"}[type(solver)]


def require_b200(solver: AbstractSolvers):
    if not isinstance(solver, b200_solver_def):
        raise NotImplementedError(
            f"mpc_solver resolves to {solver_name(solver)}: that is the reference's CPU path (JuMP + {solver_name(solver)}); this package "
            "implements only mpc_solver='b200' and has no CPU fallback")
