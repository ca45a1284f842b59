"""Import alias for the package directory `oceantransportmatrixbuilder.jl_b200/` (its name has a
dot, so it cannot be imported by name).  `import otmb_b200` gives the package itself."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "oceantransportmatrixbuilder.jl_b200"
_spec = importlib.util.spec_from_file_location("otmb_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["otmb_b200"] = _mod
_spec.loader.exec_module(_mod)
