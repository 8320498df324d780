"""``qmps.exact_loschmidt`` -- drop-in name for ``qmps_b200.exact_loschmidt`` (same signatures as the reference module)."""
from qmps_b200.exact_loschmidt import *  # noqa: F401,F403
from qmps_b200.exact_loschmidt import __all__  # noqa: F401
