"""Import helper: the package directory is named `ei-keyword-spotting_b200` (not a valid Python identifier),
so it is loaded by path and registered as `eikws_b200`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_ROOT, "ei-keyword-spotting_b200")


def load():
    if "eikws_b200" in sys.modules:
        return sys.modules["eikws_b200"]
    spec = importlib.util.spec_from_file_location("eikws_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["eikws_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
