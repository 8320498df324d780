from qmps_b200.loschmidts.exact_loschmidt import *  # noqa: F401,F403
