"""Alias of ``comfyui-float_optimized_b200/synth.py`` (the seeded synthetic weights / inputs moved into the package so that the
product arm of ``bench.py`` imports nothing from ``oracle/``); kept so that oracle-side tooling can keep importing it here."""
import importlib.util
import os
import sys

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "comfyui-float_optimized_b200", "synth.py")
_NAME = "float_fmt_b200_synth"
if _NAME in sys.modules:
    _m = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(_NAME, _PATH)
    _m = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _m
    _spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
