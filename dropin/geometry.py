"""`import geometry` (preprocess.py:15) resolves here."""
from opticalflowfromdepth_b200.geometry import *  # noqa: F401,F403
from opticalflowfromdepth_b200.geometry import __all__  # noqa: F401
