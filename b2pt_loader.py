"""Imports the package directory (whose name is not a Python identifier) under the name ``b2pt``."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "final-project-monte-carlo-path-tracer-with-microfacet-bsdf_b200")


def load():
    if "b2pt" in sys.modules:
        return sys.modules["b2pt"]
    spec = importlib.util.spec_from_file_location("b2pt", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2pt"] = mod
    spec.loader.exec_module(mod)
    return mod
