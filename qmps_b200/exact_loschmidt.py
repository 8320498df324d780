"""Drop-in mirror of ``qmps.exact_loschmidt`` / ``qmps.loschmidts.exact_loschmidt``:
the analytic TFIM Loschmidt rate function (qmps/loschmidts/exact_loschmidt.py:6-23),
evaluated for whole time grids by one kernel launch (Gauss-Legendre quadrature of
-(1/pi) int_0^pi log|cos^2 phi_k + sin^2 phi_k exp(-2 i t eps_k)| dk).
"""
import numpy as np

from . import batched

__all__ = ["f", "loschmidt", "loschmidts"]


def loschmidts(T, g0, g1):
    return batched.loschmidt_rate(np.asarray(T, dtype=np.float64), g0, g1).cpu().numpy()


def loschmidt(t, g0, g1):
    return float(loschmidts([t], g0, g1)[0])


def f(z, g0, g1):
    """Half of the rate function: the reference's ``f(it) + f(-it)`` is twice the real part
    of ``f(it)``; only imaginary arguments z = +-i t occur on the path."""
    z = complex(z)
    if abs(z.real) > 0:
        raise NotImplementedError("f(z) is evaluated on the GPU for purely imaginary z only")
    return 0.5 * loschmidt(z.imag, g0, g1)
