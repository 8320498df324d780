"""``qmps.ground_state`` -- drop-in name for ``qmps_b200.ground_state`` (same signatures as the reference module)."""
from qmps_b200.ground_state import *  # noqa: F401,F403
from qmps_b200.ground_state import __all__  # noqa: F401
