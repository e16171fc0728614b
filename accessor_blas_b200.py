"""Import shim: the package directory is `accessor-blas_b200/` (hyphenated, as
the project is named); this module makes it importable as
`accessor_blas_b200`."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent / "accessor-blas_b200"
_spec = importlib.util.spec_from_file_location(
    "accessor_blas_b200", _pkg_dir / "__init__.py",
    submodule_search_locations=[str(_pkg_dir)])
_module = importlib.util.module_from_spec(_spec)
sys.modules["accessor_blas_b200"] = _module
_spec.loader.exec_module(_module)
