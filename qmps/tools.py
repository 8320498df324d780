"""``qmps.tools`` -- drop-in name for ``qmps_b200.tools`` (same signatures as the reference module)."""
from qmps_b200.tools import *  # noqa: F401,F403
from qmps_b200.tools import __all__  # noqa: F401
