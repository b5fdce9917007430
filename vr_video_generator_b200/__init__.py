"""Importable alias of the package directory `vr-video-generator_b200/` (a hyphen is not a legal
module name).  Importing `vr_video_generator_b200` executes `vr-video-generator_b200/__init__.py`
under this name, with submodules resolved from that directory."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vr-video-generator_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_real, "__init__.py"), submodule_search_locations=[_real])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
