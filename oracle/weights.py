"""Synthetic weights/audio generators live in the package (pure data generation, shared with bench.py);
re-exported here so oracle users keep one import site.  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "retrieval-based-voice-conversion-webui_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from rvc_b200.synthetic import *  # noqa: F401,F403,E402
from rvc_b200.synthetic import HUBERT_CONV, V1_32K_CONFIG, V1_40K_CONFIG, V1_48K_CONFIG, V2_32K_CONFIG, V2_48K_CONFIG, _r16  # noqa: F401,E402
