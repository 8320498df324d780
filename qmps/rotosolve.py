"""``qmps.rotosolve`` -- drop-in name for ``qmps_b200.rotosolve`` (same signatures as the reference module)."""
from qmps_b200.rotosolve import *  # noqa: F401,F403
from qmps_b200.rotosolve import __all__  # noqa: F401
