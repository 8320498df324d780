"""``qmps.represent`` -- drop-in name for ``qmps_b200.represent`` (same signatures as the reference module)."""
from qmps_b200.represent import *  # noqa: F401,F403
from qmps_b200.represent import __all__  # noqa: F401
