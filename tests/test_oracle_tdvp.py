"""Classical iTDVP: the oracle restatement (oracle/tdvp.py) pinned by what the reference's call
sites and plots force -- xmps itself is not vendored, so there are no recorded vectors -- and the device algorithm
(qmps_b200/csrc/tdvp.cuh, compiled for the host by tests/host_emu) against the oracle."""
import ctypes
import os

import numpy as np
import pytest
from scipy.linalg import expm
from scipy.stats import unitary_group

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
X = np.array([[0, 1], [1, 0.0]])
I2 = np.eye(2)


def lc(D, seed):
    return O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=seed))


def ground_state(D, g, steps=600, dt=0.05, seed=0):
    """imaginary-time TDVP from a random state (what xmps' find_ground_state does in spirit)."""
    A = lc(D, seed)
    return O.tdvp_trajectory(A, O.tfim_matrix(g), dt, steps, method="euler", imaginary=True)[-1]


@pytest.mark.parametrize("D", [2, 3, 4, 5])
def test_tangent_vector_properties(D):
    A = lc(D, 30 + D)
    h = O.tfim_matrix(0.7)
    dA, e = O.dA_dt(A, h)
    assert abs(e - O.energy_transfer(A, h)) < 1e-12                  # the energy the ground-state costs use
    assert np.abs(np.einsum("ski,skj->ij", A.conj(), dA)).max() < 1e-12      # left gauge condition
    # gauge covariance: A' = c g A g^-1  =>  dA' = c g dA g^-1
    rng = np.random.default_rng(D)
    g = rng.normal(size=(D, D)) + 1j * rng.normal(size=(D, D))
    gi = np.linalg.inv(g)
    dAg, eg = O.dA_dt(1.7 * np.einsum("ab,sbc,cd->sad", g, A, gi), h)
    assert np.abs(dAg - 1.7 * np.einsum("ab,sbc,cd->sad", g, dA, gi)).max() < 1e-10 and abs(eg - e) < 1e-12
    # the flow is norm preserving to first order: d/dt <psi|psi> = 0  <=>  Re sum tr(A^dagger dA r) = 0
    _, _, r = O.eigs(A)
    assert abs(np.einsum("sij,sik,kj->", A.conj(), dA, r.T.conj().T)) < 1e-10


def test_rk4_conserves_energy_and_integrates_single_site_dynamics_exactly():
    A = lc(4, 3)
    h = O.tfim_matrix(0.7)
    traj = O.tdvp_trajectory(A, h, 0.01, 20)
    es = [O.energy_density(a, h) for a in traj]
    assert max(es) - min(es) < 1e-8
    for a in traj:
        assert O.is_left_canonical(a)
    h1 = (np.kron(X, I2) + np.kron(I2, X)) / 2          # sum_n X_n: the exact evolution is a local rotation
    traj = O.tdvp_trajectory(A, h1, 0.01, 50)
    At = np.tensordot(expm(-1j * X * 0.5), A, [1, 0])
    assert abs(O.overlap(traj[-1], At) - 1) < 1e-10
    # Euler agrees with RK4 to first order
    e1 = O.tdvp_trajectory(A, h, 1e-3, 10, method="euler")[-1]
    r1 = O.tdvp_trajectory(A, h, 1e-3, 10)[-1]
    assert abs(O.overlap(e1, r1) - 1) < 1e-6


def test_imaginary_time_flow_reaches_the_d2_ground_state():
    """TFIM g = 1, D = 2: the same optimum the energy minimisation finds (tests/test_oracle.py: -1.27254249,
    below the reference's D2_gse = -1.269909412573 and above the exact E0)."""
    A = ground_state(2, 1.0, steps=1500, dt=0.05, seed=1)
    e = O.energy_density(A, O.tfim_matrix(1.0))
    assert abs(e - (-1.27254249)) < 2e-5 and e > O.tfim_e0_exact(1.0)


def test_quench_loschmidt_rate_follows_the_analytic_tfim_curve(golden):
    """qmps/loschmidts/mps_loschmidts.py plots Trajectory(...).loschmidts() over the analytic rate function of
    qmps/loschmidts/exact_loschmidt.py.  D = 4, g0 = 1.5 -> g1 = 0.2 (scripts/loschmidt.py:336): the TDVP rate
    reproduces the analytic curve at short times (here to 1 % up to t = 0.5, before the first cusp)."""
    g0, g1, dt = 1.5, 0.2, 0.01
    A0 = ground_state(4, g0, steps=800, dt=0.05)
    assert abs(O.energy_density(A0, O.tfim_matrix(g0)) - O.tfim_e0_exact(g0)) < 1e-4
    traj = O.tdvp_trajectory(A0, O.tfim_matrix(g1), dt, 50)
    ls = O.loschmidt_rates(traj)
    assert ls[0] < 1e-12
    for k in (10, 25, 50):
        exact = float(O.exact_loschmidt(k * dt, g0, g1))
        assert abs(ls[k] - exact) < 0.01 * exact
    ref = golden["ref_exact_loschmidt"]                    # the reference's own function at t = 0.5
    k = int(np.argmin(np.abs(ref["t"] - 0.5)))
    assert abs(ref["t"][k] - 0.5) < 1e-12 and abs(ls[50] - ref["g15_02"][k]) < 0.01 * ref["g15_02"][k]


@pytest.fixture(scope="module")
def emu(built):
    return ctypes.CDLL(os.path.join(ROOT, "tests", "host_emu", "libqmps_emu.so"))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("d,D", [(2, 2), (2, 3), (2, 4), (2, 5), (2, 8), (3, 3)])
def test_device_tangent_algorithm_matches_oracle(emu, d, D):
    """tdvp.cuh (the code the CUDA kernel runs, one lane on the host) against oracle.tdvp: left-canonical tangent
    vector, energy, imaginary-time variant, and the gauge-back transform used for tensors in a general gauge."""
    N = 3
    rng = np.random.default_rng(100 * d + D)
    A = np.stack([np.linalg.qr(rng.normal(size=(d * D, D)) + 1j * rng.normal(size=(d * D, D)))[0].reshape(D, d, D).transpose(1, 0, 2)
                  for _ in range(N)])
    A = np.ascontiguousarray(A)
    hm = rng.normal(size=(d * d, d * d)) + 1j * rng.normal(size=(d * d, d * d))
    hm = np.ascontiguousarray(hm + hm.conj().T)
    for imag in (0, 1):
        out = np.zeros_like(A); en = np.zeros(N); st = np.zeros(N, np.int32)
        assert emu.emu_tdvp_tangent(d, D, ctypes.c_int64(N), P(A), P(hm), imag, P(out), P(en), P(st)) == 0
        assert not st.any()
        for k in range(N):
            dA0, e0 = O.tdvp_tangent_left_canonical(A[k], hm, imaginary=bool(imag))
            assert np.abs(out[k] - dA0).max() < 1e-9 and abs(en[k] - e0) < 1e-11
    # gauge back: scale * L^-1 B L with L upper triangular
    Lm = np.ascontiguousarray(np.stack([np.triu(rng.normal(size=(D, D)) + 1j * rng.normal(size=(D, D))) + 2 * np.eye(D) for _ in range(N)]))
    scale = np.array([1.3, 0.7, 1.0])
    back = np.zeros_like(A)
    emu.emu_gauge_back(d, D, ctypes.c_int64(N), P(A), P(Lm), P(scale), P(back))
    for k in range(N):
        ref = scale[k] * np.einsum("ab,sbc,cd->sad", np.linalg.inv(Lm[k]), A[k], Lm[k])
        assert np.abs(back[k] - ref).max() < 1e-11
