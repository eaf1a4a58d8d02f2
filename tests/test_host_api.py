"""CPU tests of the host mirror and of the C-ABI boundary (no compute calls: there is no GPU here)."""
import ctypes
import pathlib
import re

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(mpc):
    hdr = (ROOT / "include" / "mpcb200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mpcb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 27
    L = mpc._lib.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(mpc._lib.EXPORTED_SYMBOLS) == declared          # the Python binding tracks the header
    assert L.mpcb_version() == 100


def test_struct_layouts_match_header(mpc):
    """ctypes mirrors must have the C layout (sizes computed by hand from include/mpcb200.h)."""
    L = mpc._lib
    assert ctypes.sizeof(L.Settings) == 7 * 8 + 16 * 4
    assert ctypes.sizeof(L.LinearDesc) == 3 * 4 + 4 + 10 * 8 + 2 * 4
    assert ctypes.sizeof(L.Info) == 10 * 4 + 3 * 8
    assert ctypes.sizeof(L.BatchIO) == 8 + 3 * 8 + 2 * 4 + 14 * 8
    assert ctypes.sizeof(L.NnDesc) == 6 * 4 + 4 * 8
    assert ctypes.sizeof(L.NmpcDesc) == 8 + 4 + 4 + 8 * 8 + 4 + 4 + 2 * 8
    assert ctypes.sizeof(L.NmpcSettings) == ctypes.sizeof(L.Settings) + 3 * 8 + 2 * 4
    n = L.default_nmpc_settings()
    assert (n.qp.eps_abs, n.qp.eps_rel, n.qp.check_every, n.qp.sigma, n.sqp_tol, n.ls_armijo, n.ls_noise, n.sqp_max_iter, n.ls_max_halvings) == (1e-9, 0.0, 5, 0.0, 1e-6, 1e-4, 1e-10, 20, 12)
    s = L.default_settings()
    assert (s.eps_abs, s.eps_rel, s.sigma, s.alpha, s.max_iter, s.check_every, s.rho_eq_scale) == (1e-3, 1e-3, 1e-6, 1.6, 4000, 25, 1e3)


def test_no_cpu_fallback_fails_loudly(mpc, qt):
    """Without a B200 the product path must raise, never route through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(qt["xmin"], qt["xmax"]), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    with pytest.raises(mpc.MpcbError, match="no CUDA device|no CPU fallback"):
        mpc.proceed_controller(sys_, "model_predictive_control", 20, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver="b200")
    # and the package never imports the oracle
    import sys
    pkg_files = list((ROOT / "automationlabsmodelpredictivecontrol.jl_b200").glob("*.py"))
    for f in pkg_files:
        assert "oracle" not in re.sub(r'""".*?"""', "", f.read_text(), flags=re.S).replace("# the oracle", ""), f


def test_solver_table_and_auto_resolution(mpc):
    """src/sub/solver_selection.jl:9-87: legality matrix and `auto` (linear -> SCIP, non-linear -> Ipopt, MILP -> SCIP)."""
    t = mpc._IMPLEMENTATION_SOLVER_LIST
    assert set(t) == {"osqp", "scip", "ipopt", "auto", "b200"}
    lin, nl, milp = mpc.LinearProgramming(), mpc.NonLinearProgramming(), mpc.MixedIntegerLinearProgramming()
    assert mpc.solver_name(mpc.resolve_solver(lin, t["auto"])) == "SCIP"        # test/design_mpc_implementation_test.jl:84
    assert mpc.solver_name(mpc.resolve_solver(nl, t["auto"])) == "Ipopt"        # :177
    assert mpc.solver_name(mpc.resolve_solver(milp, t["auto"])) == "SCIP"
    assert mpc.solver_name(mpc.resolve_solver(lin, t["osqp"])) == "OSQP"
    assert mpc.solver_name(mpc.resolve_solver(lin, t["b200"])) == "B200" and mpc.solver_name(mpc.resolve_solver(nl, t["b200"])) == "B200"
    with pytest.raises(TypeError):
        mpc.resolve_solver(nl, t["osqp"])            # no _JuMP_model_definition(::NonLinearProgramming, ::osqp_solver_def)
    with pytest.raises(TypeError):
        mpc.resolve_solver(milp, t["b200"])


def test_cpu_solver_tags_are_refused_not_emulated(mpc, qt):
    sys_ = mpc.ConstrainedLinearControlDiscreteSystem(qt["A"], qt["B"], mpc.Hyperrectangle(qt["xmin"], qt["xmax"]), mpc.Hyperrectangle(qt["umin"], qt["umax"]))
    for solver in ("auto", "osqp", "scip"):
        with pytest.raises(NotImplementedError, match="no CPU fallback"):
            mpc.proceed_controller(sys_, "model_predictive_control", 5, 5, list(qt["x_ref"]), list(qt["u_ref"]), mpc_solver=solver)


def test_reference_broadcast_and_defaults(mpc):
    r = mpc._design_reference_mpc([0.65] * 4, [1.2, 1.2], 5)         # main_mpc.jl:105-117
    assert r.x.shape == (4, 6) and r.u.shape == (2, 5) and (r.x == 0.65).all() and (r.u == 1.2).all()
    d = mpc._DEFAULT_PARAMETERS_MODEL_PREDICTIVE_CONTROL                 # main_mpc.jl:87-94
    assert d == dict(mpc_solver="auto", mpc_terminal_ingredient="none", mpc_Q=100.0, mpc_R=0.1, mpc_S=0.0, mpc_max_time=30.0)
    assert set(mpc.IMPLEMENTATION_PROGRAMMING_LIST) == {"linear", "non_linear", "mixed_linear", "fuzzy_linear"}   # types.jl:229-234
    assert mpc.proceed_controller(None, "economic_model_predictive_control", 5, 5, [0.0], [0.0]) is None          # main_mpc.jl:54-83 is commented out


def test_controller_type_fields(mpc):
    import dataclasses
    f = lambda c: [x.name for x in dataclasses.fields(c)]
    assert f(mpc.ModelPredictiveControlTuning) == ["modeler", "reference", "horizon", "weights", "terminal_ingredient", "sample_time", "max_time"]
    assert f(mpc.ModelPredictiveControlResults) == ["x", "e_x", "u", "e_u"]
    assert f(mpc.ModelPredictiveControlController) == ["system", "tuning", "initialization", "computation_results"]
    assert f(mpc.TerminalIngredient) == ["Xf", "P"] and f(mpc.WeightsCoefficient) == ["Q", "R", "S"] and f(mpc.ReferencesStateInput) == ["x", "u"]


def test_dare_rejects_bad_input(mpc):
    with pytest.raises(mpc.MpcbError):
        mpc.dare(np.eye(2) * 2.0, np.zeros((2, 1)), np.eye(2), np.eye(1))     # unstable and uncontrollable: no stabilising solution


def test_settings_fields_agree_between_header_python_and_julia(mpc):
    """mpcb_settings is passed by pointer across three languages: the field order of include/mpcb200.h, of the ctypes mirror and of the Julia
    shim must be the same list (a size check alone would not notice two int32 fields swapped)."""
    import pathlib, re
    from almpc_b200 import _lib as L
    root = pathlib.Path(__file__).resolve().parents[1]
    hdr = (root / "include" / "mpcb200.h").read_text()
    body = hdr[hdr.index("typedef struct {", hdr.index("Solver settings.")):hdr.index("} mpcb_settings;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = re.findall(r"(?:double|int32_t)\s+(\w+)(?:\[\d+\])?\s*;", body)
    py_fields = [n for n, _ in L.Settings._fields_]
    assert c_fields == py_fields, (c_fields, py_fields)
    jl = (root / "automationlabsmodelpredictivecontrol.jl_b200" / "julia" / "B200ModelPredictiveControl.jl").read_text()
    jbody = jl[jl.index("struct MpcbSettings"):jl.index("end", jl.index("struct MpcbSettings"))]
    jl_fields = re.findall(r"(\w+)::(?:Cdouble|Int32|NTuple)", jbody)
    assert jl_fields == c_fields, (jl_fields, c_fields)
    s = L.default_settings()
    assert s.cold_init == 0 and s.ladder_iter == 0 and s.n_devices == 0      # opt-in features are off in the defaults
