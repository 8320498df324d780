"""PXP scar dynamics on the device -- the mirror of the reference's ``scars.py`` (repo root): the
two-site-unit-cell TDVP step cost built on ``Map(merge(A1,A2), merge(A1',A2')).right_fixed_point()`` and the
``simulate_scars`` loop around it.  Names and argument meaning follow the reference; the cirq circuit classes
(``ScarsAnsatz``, ``ScarGate``) are not needed because the read-out is evaluated as a tensor contraction
(``csrc/kernels_scars.cuh``), which the oracle checks against the gate-by-gate circuit.
"""
import numpy as np
import torch
from scipy.linalg import expm

from . import _lib as L
from .batched import _p, _stream

__all__ = ["A", "H", "W", "scars_costs", "scars_time_evolve_cost_function", "scars_cost_fun_alternate",
           "simulate_scars", "func_list"]

_P = np.array([[0, 0], [0, 1]], dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_n = np.array([[1, 0], [0, 0]], dtype=complex)
_I = np.eye(2, dtype=complex)


def _mt(*ops):
    out = np.eye(1, dtype=complex)
    for o in ops:
        out = np.kron(out, o)
    return out


def A(theta, phi):
    """The D = 2 PXP ansatz tensor [s][i][j] (scars.py:70-73)."""
    return np.array([[[0, 1j * np.exp(-1j * phi)], [0, 0]], [[np.cos(theta), 0], [np.sin(theta), 0]]], dtype=complex)


def H(mu):
    """PXP + chemical potential on four sites, 16 x 16 (scars.py:23-27)."""
    return 0.5 * (_mt(_I, _P, _X, _P) + _mt(_P, _X, _P, _I)) + (mu / 4) * (
        _mt(_I, _I, _I, _n) + _mt(_I, _I, _n, _I) + _mt(_I, _n, _I, _I) + _mt(_n, _I, _I, _I))


def W(mu, dt):
    """The evolution gate ``expm(1j * dt * H(mu))`` as a 16 x 16 matrix (scars.py:29 wraps it in a cirq ``Tensor``)."""
    return expm(1j * dt * H(mu))


def _rdt(dtype):
    return torch.float64 if dtype == torch.complex128 else torch.float32


def scars_costs(params, current_params, ham, dtype=torch.complex128, want_eta=False):
    """Batched step cost: ``params[N, 4]`` candidates ``[theta1, phi1, phi2, theta2]`` against ``current_params``
    (``[4]`` shared or ``[N, 4]``), ``ham`` the 16 x 16 gate.  Returns ``cost[N]`` (device), optionally ``eta[N]``."""
    dev = torch.device("cuda", torch.cuda.current_device())
    p = torch.as_tensor(np.asarray(params, dtype=np.float64) if not isinstance(params, torch.Tensor) else params).to(dev, torch.float64).reshape(-1, 4).contiguous()
    c = torch.as_tensor(np.asarray(current_params, dtype=np.float64) if not isinstance(current_params, torch.Tensor) else current_params).to(dev, torch.float64).reshape(-1, 4).contiguous()
    N, NC = p.shape[0], c.shape[0]
    if NC not in (1, N):
        raise ValueError("current_params must hold one parameter set or one per candidate")
    Wd = torch.as_tensor(np.asarray(ham)).to(dev, dtype).contiguous() if not isinstance(ham, torch.Tensor) else ham.to(dev, dtype).contiguous()
    cost = torch.empty((N,), dtype=_rdt(dtype), device=dev)
    eta = torch.empty((N,), dtype=dtype, device=dev) if want_eta else None
    L.check(L.require_device().qmps_scars_cost(N, _p(p), NC, _p(c), _p(Wd), _p(cost), _p(eta), None,
                                               L.C128 if dtype == torch.complex128 else L.C64, _stream()), "scars_cost")
    return (cost, eta) if want_eta else cost


def scars_cost_fun_alternate(params, current_params, ham):
    """scars.py:113-155 for one parameter vector (``ham``: the 16 x 16 matrix of ``W(mu, dt)``)."""
    return float(scars_costs(np.asarray(params)[None], current_params, ham).cpu()[0])


# the circuit-parameterised variant (scars.py:76-111) builds the same isometries from ScarGate and returns the same number
scars_time_evolve_cost_function = scars_cost_fun_alternate


def simulate_scars(dt, timesteps, mu, initial_params, save_file=None, n_gen=6, npop=1024, sigma0=0.05, seed=0, n_bfgs=20,
                   dtype=torch.complex128, return_costs=False):
    """scars.py:157-170: ``timesteps`` optimised steps from ``initial_params``; the reference calls scipy's Nelder-Mead
    once per step, here every step is a population search + BFGS on the device (``qmps_scars_trajectory``) and the
    trajectory comes back once.  Returns the parameters BEFORE each step, mod 2 pi, like the reference."""
    dev = torch.device("cuda", torch.cuda.current_device())
    Wd = torch.as_tensor(W(mu, dt)).to(dev, dtype).contiguous()
    p0 = torch.as_tensor(np.asarray(initial_params, dtype=np.float64)).to(dev).contiguous()
    traj = torch.empty((timesteps + 1, 4), dtype=torch.float64, device=dev)
    costs = torch.empty((timesteps,), dtype=_rdt(dtype), device=dev)
    L.check(L.require_device().qmps_scars_trajectory(_p(p0), _p(Wd), int(timesteps), int(n_gen), int(npop), float(sigma0), int(seed),
                                                     int(n_bfgs), _p(traj), _p(costs),
                                                     L.C128 if dtype == torch.complex128 else L.C64, _stream()), "scars_trajectory")
    out = np.mod(traj[:-1].cpu().numpy(), 2 * np.pi)
    if save_file:
        np.save(save_file, out)
    return (out, traj.cpu().numpy(), costs.cpu().numpy()) if return_costs else out


def func_list(angles, t, mu):
    """The classical TDVP equations of motion the reference integrates with ``odeint`` (scars.py:175-181)."""
    from numpy import sin, cos, tan

    def dth(t1, p1, p2, t2):
        return tan(t2) * sin(t1) * (cos(t1) ** 2) * cos(p1) + cos(t2) * cos(p2)

    def dph(t1, p1, p2, t2):
        return 2 * tan(t1) * cos(t2) * sin(p2) - 0.5 * tan(t2) * cos(t1) * sin(p1) * (2 * (sin(t2) ** -2) + cos(2 * t1) - 5)
    a = list(angles)
    return [dth(*a), -mu + dph(*a), -mu + dph(*reversed(a)), dth(*reversed(a))]
