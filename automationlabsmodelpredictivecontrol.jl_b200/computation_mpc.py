"""The hot path, mirroring /root/reference/src/main/computation_mpc.jl:17-55, plus the batched overloads.

Single problem (reference semantics):   update_initialization(C, x0);  calculate(C)  -> C.computation_results
Batched (new):  update_initialization(C, X0[batch, nx], references=(xref, uref));  calculate(C) -> C.batch_results
"""
from __future__ import annotations

import numpy as np

from .types import ModelPredictiveControlController


def update_initialization(C: ModelPredictiveControlController, initialization, references=None):
    """computation_mpc.jl:17-29: store the state measurement; the JuMP.fix of x[:,1] becomes the x0 operand of the
    next batched solve.  `initialization` may be a vector (one problem) or a (batch, nx) matrix."""
    x0 = np.asarray(initialization, np.float64)
    m = C.tuning.modeler
    if x0.ndim == 1:
        C.initialization = x0
        m.x0 = x0[None, :]
    else:
        C.initialization = x0[0]
        m.x0 = x0
    if references is not None:
        m.xref, m.uref = np.asarray(references[0], np.float64), np.asarray(references[1], np.float64)
    else:
        m.xref, m.uref = C.tuning.reference.x[:, 0].copy(), C.tuning.reference.u[:, 0].copy()


def calculate(C: ModelPredictiveControlController, warm_start: bool = False, want=("u", "e_u", "x", "e_x", "u0", "objective")):
    """computation_mpc.jl:38-55: solve and copy u, e_u, x, e_x into `computation_results` (first problem of the batch);
    the whole batch, with per-problem status / iterations / residuals / objective, lands in C.batch_results.
    Like the reference, no termination status is checked here."""
    m = C.tuning.modeler
    if m.x0 is None:
        raise RuntimeError("calculate!: call update_initialization! first")
    warm = m.warm if (warm_start and m.warm is not None and m.warm[0].shape[0] == m.x0.shape[0]) else None
    w = tuple(want) + (tuple(k for k in ("u", "y") if k not in want) if warm_start else ())
    res = m.solve_batch(m.x0, m.xref, m.uref, want=w, warm=warm)
    if warm_start:
        m.warm = (res["u"], res["y"])
    r = C.computation_results
    if "u" in res: r.u[:, :] = res["u"][0].T
    if "e_u" in res: r.e_u[:, :] = res["e_u"][0].T
    if "x" in res: r.x[:, :] = res["x"][0].T
    if "e_x" in res: r.e_x[:, :] = res["e_x"][0].T
    C.batch_results = res
    return res


# Julia spells these with a bang; keep importable aliases that read like the reference's tests.
update_initialization_b = update_initialization
calculate_b = calculate
