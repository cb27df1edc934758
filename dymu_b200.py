"""Loader for the package directory ``planning-path_planning_b200`` (hyphenated, so not
importable by name).  ``load()`` returns it as module ``planning_path_planning_b200``."""
import importlib.util
import os
import sys

_NAME = "planning_path_planning_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "planning-path_planning_b200")


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
