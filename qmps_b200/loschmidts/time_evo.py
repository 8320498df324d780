"""Classical content of ``qmps.loschmidts.time_evo`` (qmps/loschmidts/time_evo.py:75-153,
= scripts/loschmidt.py:209-239 = qmps/new_time_evolve.py:193-221).

``obj(p, A, WW)`` is the TDVP-step / Loschmidt cost: the reference builds a six-qubit
circuit whose |0...0> amplitude is x tr(l^dagger r)/2 with (x, r) the leading right
eigenpair of Map(WW . merge(A,A), merge(B,B)), and the scripts set l := r, so the cost
is -sqrt|x|.
"""
import numpy as np

from .. import batched
from ..represent import ShallowFullStateTensor
from ..time_evolve_tools import (merge, put_env_on_left_site, put_env_on_right_site,  # noqa: F401
                                 get_env_off_left_site, get_env_off_right_site)

__all__ = ["gate", "obj", "obj_batched", "merge"]


def gate(v, symbol="U"):
    return ShallowFullStateTensor(2, v, symbol)


def obj(p, A, WW, gate=gate):
    """Cost of one parameter vector against the state tensor A evolved by the two-site gate WW."""
    g = gate(np.asarray(p, dtype=np.float64))
    cost, _, _ = batched.loschmidt_costs(g.program(), np.asarray(p, dtype=np.float64)[None],
                                         np.asarray(A, dtype=np.complex128), np.asarray(WW, dtype=np.complex128))
    return float(cost.cpu()[0, 0])


def obj_batched(ps, A, WWs, gate=gate):
    """cost[p, k] for parameter sets ps[NP, P] and gates WWs[NT, 4, 4] in one pipeline."""
    ps = np.asarray(ps, dtype=np.float64)
    g = gate(ps[0])
    cost, echo, _ = batched.loschmidt_costs(g.program(), ps, np.asarray(A, dtype=np.complex128),
                                            np.asarray(WWs, dtype=np.complex128))
    return cost, echo
