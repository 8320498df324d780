"""``qmps.loschmidts.exact_loschmidt`` (qmps/loschmidts/exact_loschmidt.py)."""
from ..exact_loschmidt import f, loschmidt, loschmidts  # noqa: F401
