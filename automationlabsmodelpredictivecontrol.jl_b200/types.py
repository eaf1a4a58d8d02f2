"""Controller types, field-for-field with /root/reference/src/types/types.jl (line numbers cited per class)."""
from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np


@dataclasses.dataclass
class ReferencesStateInput:              # types.jl:24-27
    x: np.ndarray                        # nx x (H+1)
    u: np.ndarray                        # nu x H


@dataclasses.dataclass
class WeightsCoefficient:                # types.jl:46-50
    Q: np.ndarray
    R: np.ndarray
    S: np.ndarray


@dataclasses.dataclass
class TerminalIngredient:                # types.jl:89-92
    Xf: str
    P: np.ndarray


@dataclasses.dataclass
class ModelPredictiveControlTuning:      # types.jl:114-122  (modeler::Any -> here a B200Modeler)
    modeler: Any
    reference: ReferencesStateInput
    horizon: int
    weights: WeightsCoefficient
    terminal_ingredient: TerminalIngredient
    sample_time: float
    max_time: int


@dataclasses.dataclass
class ModelPredictiveControlResults:     # types.jl:134-139
    x: np.ndarray                        # nx x (H+1)
    e_x: np.ndarray
    u: np.ndarray                        # nu x H
    e_u: np.ndarray


@dataclasses.dataclass
class ModelPredictiveControlController:  # types.jl:151-156
    system: Any
    tuning: ModelPredictiveControlTuning
    initialization: np.ndarray
    computation_results: ModelPredictiveControlResults


# ---- solver tags (types.jl:168-192) + the one added tag ---------------------------------------------------------
class AbstractSolvers: ...
class osqp_solver_def(AbstractSolvers): ...
class highs_solver_def(AbstractSolvers): ...
class ipopt_solver_def(AbstractSolvers): ...
class scip_solver_def(AbstractSolvers): ...
class auto_solver_def(AbstractSolvers): ...
class b200_solver_def(AbstractSolvers):
    """New tag: batched CUDA ADMM / SQP on B200 through libmpcb200."""


# ---- method tags (types.jl:205-234) -----------------------------------------------------------------------------
class AbstractImplementation: ...
class MixedIntegerLinearProgramming(AbstractImplementation): ...
class NonLinearProgramming(AbstractImplementation): ...
class LinearProgramming(AbstractImplementation): ...
class FuzzyProgramming(AbstractImplementation): ...


IMPLEMENTATION_PROGRAMMING_LIST = {      # types.jl:229-234
    "linear": LinearProgramming(),
    "non_linear": NonLinearProgramming(),
    "mixed_linear": MixedIntegerLinearProgramming(),
    "fuzzy_linear": FuzzyProgramming(),
}
