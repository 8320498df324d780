"""Two-layer brick-wall family (new_tdvp/ClassicalTDVPStripped.py).

CPU: the oracle restatement (oracle/brickwall.py) against outputs of the REFERENCE's own code
(tests/golden/ref_brickwall.npz, made by oracle/make_golden_bw.py) and against the known answers of
new_tdvp/testTDVPStripped.py; the device algorithms compiled for the host (tests/host_emu).
GPU (-m gpu): the CUDA kernels through the C ABI against the same goldens and the oracle.
"""
import ctypes
import os

import numpy as np
import pytest
from scipy.stats import unitary_group

from oracle import brickwall as OB

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
X0 = np.array([[0, 1], [1, 0]], dtype=complex)
Z0 = np.array([[1, 0], [0, -1]], dtype=complex)
I2 = np.eye(2, dtype=complex)
HAD = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)


def kron(*ms):
    out = np.array([[1.0 + 0j]])
    for m in ms:
        out = np.kron(out, m)
    return out


def dag(x):
    return np.conj(np.swapaxes(x, -1, -2))


def same_up_to_phase(a, b, tol):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    k = np.argmax(np.abs(b))
    return np.abs(a * (b[k] / a[k]) - b).max() < tol and abs(abs(b[k] / a[k]) - 1) < tol


@pytest.fixture(scope="module")
def gold(golden):
    return golden["ref_brickwall"]


# ------------------------------------------------------------------ oracle vs the reference's outputs
def test_oracle_matches_reference_outputs(gold):
    g = gold
    for k in range(len(g["U1"])):
        U1, U2, V1, V2 = g["U1"][k], g["U2"][k], g["V1"][k], g["V2"][k]
        Mr = OB.bw_right_env_matrix(U1, U2, dag(V1), dag(V2))
        Ml = OB.bw_left_env_matrix(U1, U2, dag(V1), dag(V2))
        assert np.abs(Mr - g["renv_mat"][k]).max() < 1e-13
        assert np.abs(Ml - g["lenv_mat"][k]).max() < 1e-13
        e, v = OB.bw_exact_environment(Mr)
        assert abs(e - g["renv_eta"][k]) < 1e-12 and np.abs(v - g["renv_vec"][k]).max() < 1e-11
        e, v = OB.bw_exact_environment(Ml)
        assert abs(e - g["lenv_eta"][k]) < 1e-12 and np.abs(v - g["lenv_vec"][k]).max() < 1e-11
        assert np.abs(OB.bw_right_env_circuit(U1, U2, dag(V1), dag(V2), g["M"][k]) - g["renv_circuit"][k]).max() < 1e-13
        assert np.abs(OB.bw_right_env_matrix(U1, U2, dag(U1), dag(U2)) - g["renv_mat_same"][k]).max() < 1e-13
        assert abs(OB.bw_expectation(U1, U2, g["O2"][k]) - g["exp2"][k]) < 1e-12
        assert abs(OB.bw_expectation(U1, U2, g["O4"][k]) - g["exp4"][k]) < 1e-12
        c, ov, eta, M = OB.bw_exact_cost(U1, U2, V1, V2, g["W"])
        assert abs(c - g["exact_cost"][k]) < 1e-12 and abs(ov - g["overlap"][k]) < 1e-11


def test_oracle_known_answers_of_reference_tests():
    """new_tdvp/testTDVPStripped.py:72-144 (expectation values), :147-170 (right environment),
    :173-232 (manifold overlap)."""
    II, XX, HH = kron(I2, I2), kron(X0, X0), kron(HAD, HAD)
    assert np.isclose(OB.bw_expectation(II, II, kron(Z0, Z0)), 1)
    assert np.isclose(OB.bw_expectation(XX, II, kron(Z0, Z0)), 1)
    assert np.isclose(OB.bw_expectation(XX, II, kron(I2, Z0)), -1)
    assert np.isclose(OB.bw_expectation(HH, II, kron(X0, X0)), 1)
    assert np.isclose(OB.bw_expectation(HH, XX, kron(X0, I2)), -1)
    assert np.isclose(OB.bw_expectation(II, II, kron(Z0, Z0, Z0, Z0)), 1)
    assert np.isclose(OB.bw_expectation(XX, II, kron(Z0, Z0, Z0, Z0)), 1)
    assert np.isclose(OB.bw_expectation(XX, II, kron(I2, Z0, Z0, Z0)), -1)
    assert np.isclose(OB.bw_expectation(HH, XX, kron(X0, I2, I2, I2)), -1)
    # right environment
    assert np.allclose(OB.bw_right_env_circuit(XX, II, dag(XX), II, Z0), I2)
    M = OB.bw_right_env_matrix(XX, II, dag(XX), II)
    assert np.allclose(M, np.array([[1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]]))
    eta, vec = OB.bw_exact_environment(M)
    assert np.isclose(eta, 1) and np.allclose(vec, np.eye(2) / np.sqrt(2))
    # manifold overlap: W = 1 returns the state itself
    rs = np.random.RandomState(5)
    for _ in range(4):
        U1, U2 = unitary_group.rvs(4, random_state=rs), unitary_group.rvs(4, random_state=rs)
        ov = OB.bw_manifold_overlap(U1, U2, dag(U1), dag(U2), I2, I2, np.eye(16))
        assert np.isclose(-abs(ov) ** 2, -1)
    XI = kron(X0, I2)
    for ops, want in (((Z0, I2, I2, I2), -1), ((I2, Z0, I2, I2), 1), ((Z0, I2, Z0, I2), 1), ((I2, I2, Z0, I2), -1)):
        assert np.isclose(OB.bw_manifold_overlap(XI, II, XI, II, I2, I2, kron(*ops)), want)
    assert np.isclose(OB.bw_manifold_overlap(II, XX, II, XX, I2, I2, kron(Z0, Z0, Z0, Z0)), 1)
    assert np.isclose(OB.bw_manifold_overlap(II, XX, II, XX, I2, I2, kron(Z0, Z0, Z0, I2)), -1)


# ------------------------------------------------------------------ device algorithms on the host
def _emu(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "host_emu", "libqmps_emu.so"))
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emu_brickwall(lib, mode, U1, U2, B1=None, B2=None, side=0, undag=0, mbits=0, Mr=None, Ml=None, W=None):
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.complex128)   # noqa: E731
    U1, U2, B1, B2, Mr, Ml, W = map(c, (U1, U2, B1, B2, Mr, Ml, W))
    cnt = lambda a, n: 1 if a is None else a.reshape(-1, n, n).shape[0]               # noqa: E731
    nk, nb, nm = cnt(U1, 4), cnt(B1, 4), cnt(Mr, 2)
    nw = 1 if W is None else W.reshape(-1, 4 if (mode == 2 and mbits == 2) else 16, 4 if (mode == 2 and mbits == 2) else 16).shape[0]
    N = max(nk, nb, nm, nw)
    mat = np.zeros((N, 4, 4), complex); eta = np.zeros(N, complex); vec = np.zeros((N, 2, 2), complex)
    ov = np.zeros(N, complex); real = np.zeros(N); st = np.zeros(N, np.int32)
    lib.emu_brickwall.argtypes = [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64] + [ctypes.c_void_p] * 7
    rc = lib.emu_brickwall(mode, side, undag, mbits, N, nk, _ptr(U1), _ptr(U2), nb, _ptr(B1), _ptr(B2), nm, _ptr(Mr),
                           _ptr(Ml), nw, _ptr(W), _ptr(mat), _ptr(eta), _ptr(vec), _ptr(ov), _ptr(real), _ptr(st))
    assert rc == 0
    return dict(mat=mat, eta=eta, vec=vec, overlap=ov, real=real, status=st)


def test_host_emu_of_device_algorithms_matches_reference_outputs(built, gold):
    lib, g = _emu(built), gold
    U1, U2, V1, V2 = g["U1"], g["U2"], g["V1"], g["V2"]
    r = emu_brickwall(lib, 0, U1, U2, dag(V1), dag(V2), side=0)
    assert np.abs(r["mat"] - g["renv_mat"]).max() < 1e-13
    assert np.abs(r["eta"] - g["renv_eta"]).max() < 1e-12
    assert np.abs(r["vec"] - g["renv_vec"]).max() < 1e-10
    assert not r["status"].any()
    r = emu_brickwall(lib, 0, U1, U2, V1, V2, side=1, undag=1)
    assert np.abs(r["mat"] - g["lenv_mat"]).max() < 1e-13
    assert np.abs(r["eta"] - g["lenv_eta"]).max() < 1e-12
    assert np.abs(r["vec"] - g["lenv_vec"]).max() < 1e-10
    r = emu_brickwall(lib, 1, U1, U2, dag(V1), dag(V2), Mr=g["M"])
    assert np.abs(r["vec"] - g["renv_circuit"]).max() < 1e-13
    r = emu_brickwall(lib, 2, U1, U2, mbits=2, W=g["O2"])
    assert np.abs(r["real"] - g["exp2"]).max() < 1e-12
    r = emu_brickwall(lib, 2, U1, U2, mbits=4, W=g["O4"])
    assert np.abs(r["real"] - g["exp4"]).max() < 1e-12
    Mr = g["renv_vec"]
    r = emu_brickwall(lib, 3, U1, U2, dag(V1), dag(V2), Mr=Mr, Ml=dag(Mr), W=g["W"])
    assert np.abs(r["overlap"] - g["overlap"]).max() < 1e-12
    r = emu_brickwall(lib, 4, U1, U2, V1, V2, undag=1, W=g["W"])
    assert np.abs(r["real"] - g["exact_cost"]).max() < 1e-11
    assert np.abs(r["overlap"] - g["overlap"]).max() < 1e-10
    # same-state environment: rank-one map, eigenvalue 1, eigenvector 1/sqrt(2) (testTDVPStripped.py:156-170)
    r = emu_brickwall(lib, 0, U1, U2, U1, U2, side=0, undag=1)
    assert np.abs(r["mat"] - g["renv_mat_same"]).max() < 1e-13
    assert np.abs(r["eta"] - 1).max() < 1e-12
    for k in range(len(U1)):
        assert same_up_to_phase(r["vec"][k], np.eye(2) / np.sqrt(2), 1e-9)


def test_host_emu_thread_per_candidate_cost_matches_reference_outputs(built, gold):
    """brickwall.cuh::bw_cost_thread (the thread body of bw_cost_thread_kernel: one ket state, one W,
    registers only, bra contracted along its matrix-product structure) against the reference's own
    Evolve.exact_cost_function outputs."""
    lib, g = _emu(built), gold
    lib.emu_bw_cost_thread.argtypes = [ctypes.c_int64] + [ctypes.c_void_p] * 10
    W = np.ascontiguousarray(g["W"])
    for k in range(len(g["U1"])):
        U1, U2 = np.ascontiguousarray(g["U1"][k]), np.ascontiguousarray(g["U2"][k])
        V1 = np.ascontiguousarray(np.stack([g["V1"][k], g["U1"][k]]))          # the candidate, and the state itself
        V2 = np.ascontiguousarray(np.stack([g["V2"][k], g["U2"][k]]))
        cost = np.zeros(2); ov = np.zeros(2, complex); eta = np.zeros(2, complex)
        Mr = np.zeros((2, 2, 2), complex); st = np.zeros(2, np.int32)
        assert lib.emu_bw_cost_thread(2, _ptr(U1), _ptr(U2), _ptr(V1), _ptr(V2), _ptr(W), _ptr(cost), _ptr(ov), _ptr(eta),
                                      _ptr(Mr), _ptr(st)) == 0
        assert not st.any()
        assert abs(cost[0] - g["exact_cost"][k]) < 1e-12 and abs(ov[0] - g["overlap"][k]) < 1e-11
        assert abs(eta[0] - g["renv_eta"][k]) < 1e-12 and np.abs(Mr[0] - g["renv_vec"][k]).max() < 1e-10
        c0, ov0, eta0, M0 = OB.bw_exact_cost(U1, U2, U1, U2, W)
        assert abs(cost[1] - c0) < 1e-12 and abs(eta[1] - 1) < 1e-12 and same_up_to_phase(Mr[1], np.eye(2) / np.sqrt(2), 1e-9)


def test_host_emu_thread_per_problem_environment_matches_reference_outputs(built, gold):
    """brickwall.cuh::bw_env_thread (thread body of bw_env_thread_kernel): right / left exact environments in both
    bra conventions against the reference's own outputs, and the same-state rank-one map (eta = 1, 1/sqrt 2)."""
    lib, g = _emu(built), gold
    lib.emu_bw_env_thread.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 8
    c = lambda a: np.ascontiguousarray(a, dtype=np.complex128)          # noqa: E731
    N = len(g["U1"])
    U1, U2 = c(g["U1"]), c(g["U2"])
    for side, undag, B1, B2, km, ke, kv in ((0, 0, c(dag(g["V1"])), c(dag(g["V2"])), "renv_mat", "renv_eta", "renv_vec"),
                                            (1, 1, c(g["V1"]), c(g["V2"]), "lenv_mat", "lenv_eta", "lenv_vec"),
                                            (0, 1, U1, U2, "renv_mat_same", "renv_eta_same", None)):
        mat = np.zeros((N, 4, 4), complex); eta = np.zeros(N, complex); vec = np.zeros((N, 2, 2), complex)
        st = np.zeros(N, np.int32)
        assert lib.emu_bw_env_thread(side, undag, N, _ptr(U1), _ptr(U2), _ptr(B1), _ptr(B2), _ptr(mat), _ptr(eta),
                                     _ptr(vec), _ptr(st)) == 0
        assert not st.any()
        assert np.abs(mat - g[km]).max() < 1e-13 and np.abs(eta - g[ke]).max() < 1e-12
        if kv:
            assert np.abs(vec - g[kv]).max() < 1e-10
        else:
            for k in range(N):
                assert same_up_to_phase(vec[k], np.eye(2) / np.sqrt(2), 1e-9)


# ------------------------------------------------------------------ GPU parity through the C ABI
def _haar(n, count, seed):
    rs = np.random.RandomState(seed)
    return np.stack([unitary_group.rvs(n, random_state=rs) for _ in range(count)])


@pytest.mark.gpu
def test_gpu_brickwall_matches_reference_outputs(built, gold):
    import torch
    from qmps_b200 import brickwall as BW
    g = gold
    U1, U2, V1, V2 = g["U1"], g["U2"], g["V1"], g["V2"]
    r = BW.bw_environment(U1, U2, dag(V1), dag(V2), side="right")
    torch.cuda.synchronize()
    assert np.abs(r.mat.cpu().numpy() - g["renv_mat"]).max() < 1e-13
    assert np.abs(r.eta.cpu().numpy() - g["renv_eta"]).max() < 1e-12
    assert np.abs(r.vec.cpu().numpy() - g["renv_vec"]).max() < 1e-10
    assert int(r.status.abs().sum()) == 0
    r = BW.bw_environment(U1, U2, V1, V2, side="left", bra_undaggered=True)
    assert np.abs(r.mat.cpu().numpy() - g["lenv_mat"]).max() < 1e-13
    assert np.abs(r.eta.cpu().numpy() - g["lenv_eta"]).max() < 1e-12
    assert np.abs(r.vec.cpu().numpy() - g["lenv_vec"]).max() < 1e-10
    assert np.abs(BW.bw_env_apply(U1, U2, dag(V1), dag(V2), g["M"]).cpu().numpy() - g["renv_circuit"]).max() < 1e-13
    assert np.abs(BW.bw_expectation(U1, U2, g["O2"]).cpu().numpy() - g["exp2"]).max() < 1e-12
    assert np.abs(BW.bw_expectation(U1, U2, g["O4"]).cpu().numpy() - g["exp4"]).max() < 1e-12
    Mr = g["renv_vec"]
    ov = BW.bw_overlap(U1, U2, dag(V1), dag(V2), Mr, dag(Mr), g["W"]).cpu().numpy()
    assert np.abs(ov - g["overlap"]).max() < 1e-12
    c = BW.bw_evolve_cost(U1, U2, V1, V2, g["W"], want_all=True)
    assert np.abs(c.cost.cpu().numpy() - g["exact_cost"]).max() < 1e-11
    assert np.abs(c.overlap.cpu().numpy() - g["overlap"]).max() < 1e-10
    assert np.abs(c.eta.cpu().numpy() - g["renv_eta"]).max() < 1e-12
    assert np.abs(c.Mr.cpu().numpy() - g["renv_vec"]).max() < 1e-10
    assert int(c.status.abs().sum()) == 0


@pytest.mark.gpu
def test_gpu_brickwall_batch_vs_oracle_and_broadcast(built):
    import torch
    from qmps_b200 import brickwall as BW
    N = 1000                                   # not a multiple of the problems per CTA: ragged tail
    U1, U2 = _haar(4, 1, 11)[0], _haar(4, 1, 12)[0]
    V1, V2 = _haar(4, N, 13), _haar(4, N, 14)
    from scipy.linalg import expm
    rs = np.random.RandomState(15)
    h = rs.randn(16, 16) + 1j * rs.randn(16, 16)
    W = expm(-0.05j * (h + h.conj().T))
    for k in range(0, N, 2):                   # half of the candidates close to the ket state
        a = rs.randn(4, 4) + 1j * rs.randn(4, 4)
        b = rs.randn(4, 4) + 1j * rs.randn(4, 4)
        V1[k] = U1 @ expm(0.03j * (a + a.conj().T))
        V2[k] = U2 @ expm(0.03j * (b + b.conj().T))
    c = BW.bw_evolve_cost(U1, U2, V1, V2, W, want_all=True)
    torch.cuda.synchronize()
    cost, ov, eta, Mr = (x.cpu().numpy() for x in (c.cost, c.overlap, c.eta, c.Mr))
    assert int(c.status.abs().sum()) == 0
    for k in list(range(0, 40)) + [N - 3, N - 2, N - 1]:
        c0, ov0, eta0, M0 = OB.bw_exact_cost(U1, U2, V1[k], V2[k], W)
        assert abs(cost[k] - c0) < 1e-10 * max(1, abs(c0))
        assert abs(eta[k] - eta0) < 1e-11 and np.abs(Mr[k] - M0).max() < 1e-9
        assert abs(ov[k] - ov0) < 1e-9
    # the same batch through the group kernel (thread-per-candidate path switched off)
    lib = __import__("qmps_b200._lib", fromlist=["load"]).load()
    lib.qmps_set_option(b"bw_thread", 0)
    try:
        cg = BW.bw_evolve_cost(U1, U2, V1, V2, W, want_all=True)
        torch.cuda.synchronize()
    finally:
        lib.qmps_set_option(b"bw_thread", 1)
    assert (cg.cost - c.cost).abs().max().item() < 1e-12 and (cg.overlap - c.overlap).abs().max().item() < 1e-11
    assert (cg.Mr - c.Mr).abs().max().item() < 1e-9 and int(cg.status.abs().sum()) == 0
    # complex64 mode: 1e-5
    c32 = BW.bw_evolve_cost(torch.from_numpy(U1).to(torch.complex64).cuda(), U2, V1, V2, W).cpu().numpy()
    assert c32.dtype == np.float32 and np.abs(c32 - cost).max() < 2e-5
    # W = 1 and candidate = state: the exact environment is 1/sqrt(2) (unit 2-norm, testTDVPStripped.py:169),
    # so overlap = <psi| Mr^dagger (x) 1 (x) Mr |psi> = 1/2 and the cost is -1/4 (the reference's own
    # :180-191 check passes M = eye(2) and gets -1: covered below through ManifoldOverlap.circuit)
    Us1, Us2 = _haar(4, 64, 21), _haar(4, 64, 22)
    c1 = BW.bw_evolve_cost(Us1, Us2, Us1, Us2, np.eye(16)).cpu().numpy()
    assert np.abs(c1 + 0.25).max() < 1e-10
    ov1 = BW.bw_overlap(Us1, Us2, dag(Us1), dag(Us2), np.eye(2), np.eye(2), np.eye(16)).cpu().numpy()
    assert np.abs(-np.abs(ov1) ** 2 + 1).max() < 1e-10
    # expectation values: Hermitian operator -> value inside the spectrum; identity -> 1
    e1 = BW.bw_expectation(Us1, Us2, np.eye(16)).cpu().numpy()
    assert np.abs(e1 - 1).max() < 1e-12
    # empty batch
    assert BW.bw_evolve_cost(U1, U2, V1[:0], V2[:0], W).numel() == 0


@pytest.mark.gpu
def test_gpu_brickwall_reference_classes_known_answers(built):
    """The reference's class API as a batch of one, on the known answers of testTDVPStripped.py."""
    from qmps_b200 import brickwall as BW
    r4 = lambda m: np.asarray(m).reshape(2, 2, 2, 2)       # noqa: E731
    II, XX, HH = kron(I2, I2), kron(X0, X0), kron(HAD, HAD)
    OC, RE, MO = BW.OverlapCalculator(), BW.RightEnvironment(), BW.ManifoldOverlap()
    assert np.isclose(OC.expectation_value(r4(II), r4(II), r4(kron(Z0, Z0))), 1)
    assert np.isclose(OC.expectation_value(r4(XX), r4(II), r4(kron(I2, Z0))), -1)
    assert np.isclose(OC.expectation_value(r4(HH), r4(XX), r4(kron(X0, I2))), -1)
    assert np.isclose(OC.expectation_value(r4(XX), r4(II), kron(I2, Z0, Z0, Z0).reshape((2,) * 8)), -1)
    assert np.isclose(OC.expectation_value(r4(HH), r4(XX), kron(X0, I2, I2, I2).reshape((2,) * 8)), -1)
    assert np.allclose(RE.circuit(r4(XX), r4(II), r4(dag(XX)), r4(II), Z0), I2)
    M = RE.exact_environment_circuit(r4(XX), r4(II), r4(dag(XX)), r4(II))
    assert np.allclose(M, np.array([[1, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0]]))
    eta, vec = RE.exact_environment(r4(XX), r4(II), r4(dag(XX)), r4(II))
    assert np.isclose(eta, 1) and same_up_to_phase(vec, np.eye(2) / np.sqrt(2), 1e-9)
    XI = kron(X0, I2)
    for ops, want in (((Z0, I2, I2, I2), -1), ((I2, Z0, I2, I2), 1), ((Z0, I2, Z0, I2), 1), ((I2, I2, Z0, I2), -1)):
        assert np.isclose(MO.circuit(r4(XI), r4(II), r4(XI), r4(II), I2, I2, kron(*ops).reshape((2,) * 8)), want)
    Mr, Ml = BW.Represent().exact_env(r4(XX), r4(II), r4(dag(XX)), r4(II))
    assert same_up_to_phase(Mr, np.eye(2) / np.sqrt(2), 1e-9) and Ml.shape == (2, 2)
    rs = np.random.RandomState(3)
    U1, U2 = unitary_group.rvs(4, random_state=rs), unitary_group.rvs(4, random_state=rs)
    ev = BW.Evolve(W=np.eye(16), U1=U1, U2=U2)
    assert np.isclose(ev.exact_cost_function_unitaries(U1, U2), -0.25)
