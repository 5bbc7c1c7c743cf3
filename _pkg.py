"""Registers the package directory ``learn-fhe_b200/`` (not a valid Python identifier) as ``learn_fhe_b200``."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    if "learn_fhe_b200" in sys.modules:
        return sys.modules["learn_fhe_b200"]
    pkg_dir = os.path.join(_ROOT, "learn-fhe_b200")
    spec = importlib.util.spec_from_file_location("learn_fhe_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["learn_fhe_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
