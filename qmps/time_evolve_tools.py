"""``qmps.time_evolve_tools`` -- drop-in name for ``qmps_b200.time_evolve_tools`` (same signatures as the reference module)."""
from qmps_b200.time_evolve_tools import *  # noqa: F401,F403
from qmps_b200.time_evolve_tools import __all__  # noqa: F401
