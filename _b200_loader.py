"""Imports the product package.  Its directory is named after the reference
(`montecarlo.jl_b200/`), which is not a valid Python identifier, so it is loaded under
the module name `montecarlo_jl_b200`."""
import importlib.util
import sys
from pathlib import Path

_NAME = "montecarlo_jl_b200"
_DIR = Path(__file__).resolve().parent / "montecarlo.jl_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, _DIR / "__init__.py",
                                                  submodule_search_locations=[str(_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
