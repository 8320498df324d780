from qmps_b200.loschmidts.time_evo import *  # noqa: F401,F403
