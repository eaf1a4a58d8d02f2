"""Entry point, mirroring /root/reference/src/main/main_mpc.jl."""
from __future__ import annotations

import numpy as np

from .design_mpc import _model_predictive_control_design
from .types import ReferencesStateInput


def _design_reference_mpc(state_reference, input_reference, horizon: int) -> ReferencesStateInput:
    """main_mpc.jl:105-117: constant references broadcast over the horizon."""
    xr = np.asarray(state_reference, float).reshape(-1, 1); ur = np.asarray(input_reference, float).reshape(-1, 1)
    return ReferencesStateInput(xr * np.ones((xr.shape[0], horizon + 1)), ur * np.ones((ur.shape[0], horizon)))


def proceed_controller(system, mpc_controller_type: str, mpc_horizon: int, mpc_sample_time: int, mpc_state_reference,
                       mpc_input_reference, **kws):
    """main_mpc.jl:22-84.  As in the reference, `kws` may also be passed as one mapping under the key `kws`
    (main_mpc.jl:33-34) and any controller type other than "model_predictive_control" returns nothing."""
    kws = dict(kws.get("kws", kws))
    if mpc_controller_type == "model_predictive_control":
        references = _design_reference_mpc(mpc_state_reference, mpc_input_reference, mpc_horizon)
        return _model_predictive_control_design(system, mpc_horizon, mpc_sample_time, references, **kws)
    return None
