"""Importable alias of the package directory ``pose-estimation-with-message-passing-networks_b200/``.

The directory name the build contract asks for contains hyphens, which Python cannot
import; this module loads that directory as the package ``pgmp_b200`` (sub-modules such
as ``pgmp_b200.graph_constructor`` resolve inside it).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "pose-estimation-with-message-passing-networks_b200")
_spec = importlib.util.spec_from_file_location(
    "pgmp_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["pgmp_b200"] = _mod
_spec.loader.exec_module(_mod)
