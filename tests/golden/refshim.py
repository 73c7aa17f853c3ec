"""Moved to oracle/refshim.py (the reference arm of bench.py needs it too); kept as an alias for the golden tooling."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.refshim import *  # noqa: F401,F403,E402
from oracle.refshim import REF_ROOT, build_reference_fmt, install_shims, load_reference, reference_available  # noqa: F401,E402
