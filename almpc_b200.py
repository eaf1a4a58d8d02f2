"""Import shim: the package directory is named after the reference (`automationlabsmodelpredictivecontrol.jl_b200`),
which is not a valid Python identifier, so it is loaded by path and registered under the alias `almpc_b200`."""
import importlib.util
import pathlib
import sys

_PKG_DIR = pathlib.Path(__file__).resolve().parent / "automationlabsmodelpredictivecontrol.jl_b200"
_ALIAS = "almpc_b200"

if _ALIAS not in sys.modules or getattr(sys.modules[_ALIAS], "__path__", None) is None:
    _spec = importlib.util.spec_from_file_location(_ALIAS, _PKG_DIR / "__init__.py", submodule_search_locations=[str(_PKG_DIR)])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_ALIAS] = _mod
    _spec.loader.exec_module(_mod)
