"""Import alias for the package directory ``crux.jl_b200/`` (its name is not a Python identifier).

    import crux_b200 as crux

loads ``crux.jl_b200/__init__.py`` as the package ``crux_b200`` (relative imports inside it keep working).
"""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "crux.jl_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
