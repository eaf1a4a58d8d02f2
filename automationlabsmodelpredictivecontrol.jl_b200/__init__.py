"""B200-native MPC solve loop behind the AutomationLabsModelPredictiveControl.jl controller API.

Exports mirror /root/reference/src/AutomationLabsModelPredictiveControl.jl:25-31 (`proceed_controller`,
`update_initialization!`, `calculate!`; `update_and_compute!` / `update!` are exported by the reference but their
bodies are commented out, src/main/computation_mpc.jl:58-284, so they are not provided)."""
from . import _lib
from ._lib import MpcbError, default_settings
from .computation_mpc import calculate, update_initialization
from .design_mpc import _DEFAULT_PARAMETERS_MODEL_PREDICTIVE_CONTROL, _model_predictive_control_design, dare, dare_batch
from .main_mpc import _design_reference_mpc, proceed_controller
from .modeler import B200Modeler
from .nmpc import B200NonlinearModeler
from .nn import DenseNet, Fnn, Icnn, PolyNet, Rbf, ResNet, linearize
from .solver_selection import _IMPLEMENTATION_SOLVER_LIST, resolve_solver, solver_name
from .systems import ConstrainedBlackBoxControlDiscreteSystem, ConstrainedLinearControlDiscreteSystem, Hyperrectangle
from .types import (IMPLEMENTATION_PROGRAMMING_LIST, LinearProgramming, MixedIntegerLinearProgramming,
                    ModelPredictiveControlController, ModelPredictiveControlResults, ModelPredictiveControlTuning,
                    NonLinearProgramming, ReferencesStateInput, TerminalIngredient, WeightsCoefficient, auto_solver_def,
                    b200_solver_def, ipopt_solver_def, osqp_solver_def, scip_solver_def)

__all__ = ["proceed_controller", "update_initialization", "calculate"]
