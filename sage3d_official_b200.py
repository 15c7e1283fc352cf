"""Import alias: the package directory is ``sage-3d_official_b200`` (not a valid identifier), so
``import sage3d_official_b200`` loads it from there and installs it under this name."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "sage-3d_official_b200"
_spec = importlib.util.spec_from_file_location(__name__, _dir / "__init__.py",
                                               submodule_search_locations=[str(_dir)])
_pkg = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _pkg
_spec.loader.exec_module(_pkg)
