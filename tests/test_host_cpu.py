"""CPU tests that need no GPU: the C ABI library loads and exports every symbol the header
declares, the product has no CPU fallback and never touches the oracle, the host-side
logic (gate programs, Hamiltonian, rotosolve drivers, sharding + gloo argmin), and the
device ALGORITHMS compiled for the host (tests/host_emu) against the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
from scipy.stats import unitary_group

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from qmps_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "qmps_b200.h")).read()
    declared = set(re.findall(r"\b(qmps_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 17
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.qmps_version()
    assert ctypes.sizeof(_lib.GateOp) == 32


def test_library_holds_sm100a_code_only(built):
    from qmps_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_and_no_oracle_in_product(built):
    from qmps_b200 import _lib, batched, tools
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the GPU-less container")
    with pytest.raises(_lib.QmpsError):
        batched.env_exact(U=np.eye(4)[None].astype(complex))
    with pytest.raises(_lib.QmpsError):
        tools.get_env_exact(np.eye(4, dtype=complex))
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qmps_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f
    for f in os.listdir(os.path.join(ROOT, "qmps")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "qmps", f)).read()


def test_gate_programs_have_reference_gate_order():
    from qmps_b200 import represent as R, _lib as L
    g = R.ShallowFullStateTensor(2, np.zeros(15)).program()
    codes = [op[0] for op in g.ops]
    assert codes == [L.G_RZ, L.G_RX, L.G_RZ, L.G_RZ, L.G_RX, L.G_RZ, L.G_CNOT, L.G_RY, L.G_CNOT, L.G_RY, L.G_RZ,
                     L.G_CNOT, L.G_RZ, L.G_RX, L.G_RZ, L.G_RZ, L.G_RX, L.G_RZ]          # represent.py:392-401
    assert [op[3] for op in g.ops if op[3] >= 0] == list(range(15))
    g = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
    assert g.nq == 4 and len(g) == 3 * (8 + 3)
    assert [(op[1], op[2]) for op in g.ops[8:11]] == [(2, 3), (1, 2), (0, 1)]       # reversed ladder
    assert R.ShallowCNOTStateTensor_nonuniform.params_per_iter(8) == 8
    assert R.ShallowEnvironment(4, np.zeros(4)).num_qubits() == 4
    assert R.StateGate(np.zeros(6)).program().nq == 2


def test_hamiltonian_and_host_helpers():
    from qmps_b200 import ground_state as gs, tools, rotosolve as rs
    H = gs.Hamiltonian({"ZZ": -1, "X": 1}).to_matrix()
    assert np.allclose(H, O.tfim_matrix(1.0))
    assert np.allclose(gs.Hamiltonian().from_matrix(H).to_matrix(), H)
    U = gs.SU(np.random.default_rng(0).normal(size=15), 4)
    assert np.allclose(U.conj().T @ U, np.eye(4)) and abs(np.linalg.det(U) - 1) < 1e-12
    v = np.arange(10.0)
    assert np.array_equal(tools.from_real_vector(v), O.from_real_vector(v))
    assert np.array_equal(tools.to_real_vector(U), O.to_real_vector(U))
    assert np.array_equal(tools.direct_sum(np.eye(2), np.ones((1, 1))), O.direct_sum(np.eye(2), np.ones((1, 1))))
    assert tools.split_ns(list(range(7)), 3) == [[0, 1, 2], [3, 4, 5], [6]]
    assert abs(rs.rotosolve_theta(0.3, -0.2, 0.9) - O.rotosolve_theta3(0.3, -0.2, 0.9)) < 1e-15
    assert np.allclose(rs.double_rotosolve_coefficients(1, 2, 3, 4, 5, 6), O.double_rotosolve_fit(1, 2, 3, 4, 5, 6))


def test_double_rotosolve_driver_matches_reference(golden):
    """The host driver (pure Python around a cost callable) against the reference's outputs."""
    from qmps_b200 import tools
    g = golden["ref_tools"]

    def eps(p):
        return (np.sin(p[0]) * np.cos(2 * p[1]) + 0.3 * np.sin(2 * p[0] + 0.4)
                + 0.5 * np.cos(p[1] - 0.2) + 0.1 * np.sin(p[2]) * np.sin(p[0]))
    p = g["drs_p0"].copy()
    res = tools.double_rotosolve(eps, p, 3, False)
    assert np.allclose(res.history, g["drs_history"], atol=1e-12) and np.allclose(res.x, g["drs_x"], atol=1e-12)
    assert res.x is p                                       # mutated in place like the reference
    eps.batched = lambda th: np.array([eps(t) for t in th])
    res2 = tools.double_rotosolve(eps, g["drs_p0"].copy(), 3, False)
    assert np.allclose(res2.x, g["drs_x"], atol=1e-12)


def test_shard_ranges_cover_the_batch():
    from qmps_b200.dist import shard_range
    for n in (0, 1, 7, 1 << 20, 65536 * 3 + 5):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


_GLOO_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from qmps_b200.dist import shard_range, global_argmin, global_sum
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator().manual_seed(0)
cost = torch.randn(1001, dtype=torch.float64, generator=g)
cost[700] = cost.min() - 1.0
cost[123] = cost[700]          # tie: the smaller global index must win
lo, hi = shard_range(1001)
loc = cost[lo:hi]
k = int(torch.argmin(loc))
first = int((loc == loc.min()).nonzero()[0])
best, idx = global_argmin(loc[first].reshape(1), torch.tensor([lo + first]))
assert float(best) == float(cost.min()) and int(idx) == 123, (float(best), int(idx))
s = global_sum(loc.sum().reshape(1))
assert abs(float(s) - float(cost.sum())) < 1e-9
e_best, e_idx = global_argmin(torch.tensor([0.0], dtype=torch.float64), torch.tensor([-1]))   # all shards empty
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_global_argmin_world2_gloo(tmp_path):
    script = tmp_path / "gloo_argmin.py"
    script.write_text(_GLOO_SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29617", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


# ---- the device algorithms, compiled for the host (group of one lane) -------------------------
@pytest.fixture(scope="module")
def emu(built):
    return ctypes.CDLL(os.path.join(ROOT, "tests", "host_emu", "libqmps_emu.so"))


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_emu_qr_eigenvalues(emu):
    rng = np.random.default_rng(0)
    for n in (1, 2, 4, 16, 64):
        for _ in range(5):
            M = np.ascontiguousarray((rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))) / np.sqrt(n))
            w = np.zeros(n, complex)
            assert emu.emu_eigvals(n, P(M), P(w)) == 0
            wr = np.linalg.eigvals(M)
            d = np.abs(w[:, None] - wr[None, :])
            assert max(d.min(axis=1).max(), d.min(axis=0).max()) < 1e-12
    M = np.ascontiguousarray(np.diag([1.0, 2.0, 3.0]).astype(complex))       # already triangular
    w = np.zeros(3, complex); emu.emu_eigvals(3, P(M), P(w))
    assert sorted(w.real) == [1.0, 2.0, 3.0]


def test_emu_env_d2_bloch_solver(emu):
    N = 500
    A = np.ascontiguousarray(np.stack([O.unitary_to_tensor(unitary_group.rvs(4, random_state=s)) for s in range(N)]))
    r = np.zeros((N, 2, 2), complex); eta = np.zeros(N, complex); C = np.zeros((N, 2, 2), complex); st = np.zeros(N, np.int32)
    emu.emu_env_d2(ctypes.c_int64(N), P(A), P(r), P(eta), P(C), P(st))
    assert st.sum() == 0 and np.abs(eta - 1).max() < 1e-14
    for k in range(N):
        _, r0, C0, _ = O.env_exact_parts(A[k])
        w = np.sort(np.abs(np.linalg.eigvals(O.transfer_matrix(A[k]))))[::-1]
        assert np.abs(r[k] - r0).max() < 1e-12 / (1 - w[1])
    I4 = np.ascontiguousarray(O.unitary_to_tensor(np.eye(4, dtype=complex))[None])
    emu.emu_env_d2(ctypes.c_int64(1), P(I4), P(r), P(eta), P(C), P(st))
    assert st[0] == 1                                       # product state: not positive definite


@pytest.mark.parametrize("D", [2, 4, 8])
def test_emu_generic_env_and_fixed_point(emu, D):
    N = 6
    A = np.ascontiguousarray(np.stack([O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=100 + s)) for s in range(N)]))
    B = np.ascontiguousarray(np.stack([O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=500 + s)) for s in range(N)]))
    for lc in (1, 0):
        r = np.zeros((N, D, D), complex); eta = np.zeros(N, complex); C = np.zeros((N, D, D), complex); st = np.zeros(N, np.int32)
        emu.emu_env_generic(2, D, ctypes.c_int64(N), P(A), lc, P(eta), P(r), P(C), P(st))
        assert st.sum() == 0
        for k in range(N):
            _, r0, C0, _ = O.env_exact_parts(A[k])
            assert np.abs(r[k] - r0).max() < 1e-12 and np.abs(C[k] - C0).max() < 1e-11
    for left in (0, 1):
        vec = np.zeros((N, D, D), complex); eta = np.zeros(N, complex); st = np.zeros(N, np.int32)
        emu.emu_fixed_point(2, D, ctypes.c_int64(N), P(A), P(B), left, P(eta), P(vec), P(st))
        for k in range(N):
            x, _ = (O.left_fixed_point if left else O.right_fixed_point)(A[k], B[k])
            assert abs(abs(eta[k]) - abs(x)) < 1e-12
            E = O.transfer_matrix(A[k], B[k]); Em = E.conj().T if left else E
            assert np.abs(Em @ vec[k].reshape(-1) - eta[k] * vec[k].reshape(-1)).max() < 1e-11


@pytest.mark.parametrize("D", [2, 4, 8])
def test_emu_env_real_form(emu, D):
    """envreal.cuh: the Hermitian-basis real system reproduces the oracle's environment, for
    single-site tensors (d = 2) and for two-site blocks merge(A1, A2) (d = 4)."""
    N = 6
    A = np.ascontiguousarray(np.stack([O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=700 + s)) for s in range(N)]))
    A2 = np.ascontiguousarray(np.stack([O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=800 + s)) for s in range(N)]))
    M = np.ascontiguousarray(np.stack([O.merge(A[k], A2[k]) if D == 2 else
                                       np.einsum("aik,bkj->abij", A[k], A2[k]).reshape(4, D, D) for k in range(N)]))
    for d, T in ((2, A), (4, M)):
        r = np.zeros((N, D, D), complex); st = np.zeros(N, np.int32)
        assert emu.emu_env_real(d, D, ctypes.c_int64(N), P(T), P(r), P(st)) == 0
        assert st.sum() == 0
        for k in range(N):
            _, r0, _, _ = O.env_exact_parts(T[k])
            assert np.abs(r[k] - r0).max() < 1e-12


@pytest.mark.parametrize("left", [0, 1])
def test_emu_fp_d2_register_eigenvalue_path(emu, left):
    """fp_d2.cuh: Hessenberg + QR of the 4 x 4 mixed map in registers (thread body of fp_d2_kernel)
    against numpy's eig on the oracle's transfer matrix: single-site (d = 2) and merged two-site
    (d = 4) tensors, near-identical pairs (eta close to 1) and exactly equal ones."""
    N = 40
    rs = np.random.RandomState(321)
    for d in (2, 4):
        A = np.stack([O.unitary_to_tensor(unitary_group.rvs(4, random_state=rs)) for _ in range(N)])
        B = np.stack([O.unitary_to_tensor(unitary_group.rvs(4, random_state=rs)) for _ in range(N)])
        B[:6] = A[:6]
        B[6:12] = A[6:12] + 1e-3 * (rs.randn(6, 2, 2, 2) + 1j * rs.randn(6, 2, 2, 2))
        if d == 4:
            A = np.stack([O.merge(a, a) for a in A])
            B = np.stack([O.merge(b, b) for b in B])
        A, B = np.ascontiguousarray(A), np.ascontiguousarray(B)
        eta = np.zeros(N, complex); st = np.zeros(N, np.int32); vec = np.zeros((N, 2, 2), complex)
        assert emu.emu_fp_d2(d, ctypes.c_int64(N), P(A), P(B), left, P(eta), P(st), P(vec)) == 0
        assert st.sum() == 0
        for k in range(N):
            E = O.transfer_matrix(A[k], B[k])
            Em = E.conj().T if left else E
            w = np.linalg.eigvals(Em)
            w0 = w[np.argmax(np.abs(w))]
            assert abs(abs(eta[k]) - abs(w0)) < 1e-12
            assert np.abs(w - eta[k]).min() < 1e-11          # it IS an eigenvalue of the map
            v = vec[k].reshape(-1)
            ws = np.sort(np.abs(w))[::-1]
            gap = max(np.abs(w - eta[k])[np.argsort(np.abs(w - eta[k]))[1]], 1e-3)
            assert abs(np.linalg.norm(v) - 1) < 1e-12
            assert np.abs(Em @ v - eta[k] * v).max() < 1e-11 / gap
            x0, v0 = (O.left_fixed_point if left else O.right_fixed_point)(A[k], B[k], gauge="trace")
            if abs(eta[k] - x0) < 1e-9:                    # same eigenvalue picked: same gauge-fixed vector
                assert np.abs(vec[k] - v0).max() < 1e-9 / gap


def test_emu_ansatz_and_energy(emu):
    from qmps_b200 import represent as R
    rng = np.random.default_rng(3)
    H = np.ascontiguousarray(O.tfim_matrix(1.0))
    for gate, ofn in ((R.ShallowFullStateTensor(2, rng.normal(size=15)), O.shallow_full_state_tensor),
                      (R.ShallowCNOTStateTensor_nonuniform(8, rng.normal(size=24)), lambda p: O.shallow_cnot_state_tensor_nonuniform(8, p)),
                      (R.ShallowQAOAStateTensor(4, rng.normal(size=6)), lambda p: O.shallow_qaoa_state_tensor(4, p)),
                      (R.ExactAfter4(4, rng.normal(size=12)), lambda p: O.exact_after4(4, p)),
                      (R.StateGate(rng.normal(size=6)), O.state_gate)):
        prog = gate.program(); nq = prog.nq; D = 2 ** (nq - 1)
        th = np.ascontiguousarray(gate.params[None, :])
        A = np.zeros((1, 2, D, D), complex)
        emu.emu_ansatz(prog.c_ops(), len(prog), nq, ctypes.c_int64(1), th.shape[1], P(th), 0, 1, ctypes.c_double(0.3), P(A))
        p2 = gate.params.copy(); p2[1] += 0.3
        Aref = O.unitary_to_tensor(ofn(p2))
        assert np.abs(A[0] - Aref).max() < 1e-14
        e = np.zeros(1)
        emu.emu_energy_generic(D, ctypes.c_int64(1), P(A), 0, P(H), P(e))
        assert abs(e[0] - O.energy_transfer(Aref, H)) < 1e-12
        if nq == 2:
            A2 = np.zeros((1, 2, 2, 2), complex)
            emu.emu_ansatz_reg2(prog.c_ops(), len(prog), ctypes.c_int64(1), th.shape[1], P(th), 1, ctypes.c_double(0.3), P(A2))
            assert np.abs(A2[0] - Aref).max() < 1e-14
            emu.emu_energy_d2(ctypes.c_int64(1), P(A2), P(H), P(e))
            assert abs(e[0] - O.energy_transfer(Aref, H)) < 1e-12


_PAULIS = np.ascontiguousarray(np.array([[[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], dtype=complex))


@pytest.mark.parametrize("D", [2, 4, 8])
def test_emu_canonical_forms_and_expectations(emu, D):
    """canon.cuh: left_canonicalise -> mixed -> Es on random NON-canonical tensors,
    composed exactly as the C ABI composes the device kernels (left fixed point -> gauge ->
    environment -> gauge), against oracle/canonical.py and the properties of
    tests/test_represent.py:23-31."""
    N, d = 5, 2
    rng = np.random.default_rng(40 + D)
    A = np.ascontiguousarray(rng.normal(size=(N, d, D, D)) + 1j * rng.normal(size=(N, d, D, D)))
    cN = ctypes.c_int64(N)
    lvec = np.zeros((N, D, D), complex); eta = np.zeros(N, complex); st = np.zeros(N, np.int32)
    emu.emu_fixed_point(d, D, cN, P(A), P(A), 1, P(eta), P(lvec), P(st))
    assert st.sum() == 0
    scale = np.ascontiguousarray(1 / np.sqrt(np.abs(eta)))
    AL = np.zeros_like(A); Lm = np.zeros((N, D, D), complex)
    emu.emu_gauge(d, D, cN, P(A), P(lvec), 0, P(scale), P(AL), P(Lm), P(st))
    assert st.sum() == 0
    r = np.zeros((N, D, D), complex); e1 = np.zeros(N, complex); C = np.zeros((N, D, D), complex)
    emu.emu_env_generic(d, D, cN, P(AL), 1, P(e1), P(r), P(C), P(st))
    assert st.sum() == 0
    AR = np.zeros_like(A)
    emu.emu_gauge(d, D, cN, P(AL), P(C), 1, None, P(AR), None, P(st))
    es = np.zeros((N, 3), complex)
    emu.emu_expect(d, D, cN, P(AL), P(r), None, None, 3, P(_PAULIS), P(es))
    # the general formula on the ORIGINAL tensor: right fixed point + left vector + eta
    rv = np.zeros((N, D, D), complex); eta_r = np.zeros(N, complex)
    emu.emu_fixed_point(d, D, cN, P(A), P(A), 0, P(eta_r), P(rv), P(st))
    es_gen = np.zeros((N, 3), complex)
    emu.emu_expect(d, D, cN, P(A), P(rv), P(lvec), P(eta_r), 3, P(_PAULIS), P(es_gen))
    I = np.eye(D)
    for k in range(N):
        AL0, AR0, C0 = O.mixed(A[k])
        assert np.abs(AL[k] - AL0).max() < 1e-10 and np.abs(C[k] - C0).max() < 1e-10 and np.abs(AR[k] - AR0).max() < 1e-9
        assert np.abs(np.triu(Lm[k], 0) - Lm[k]).max() == 0 and abs(np.trace(Lm[k].conj().T @ Lm[k]) - D) < 1e-10
        assert O.is_left_canonical(AL[k]) and O.is_right_canonical(AR[k], 1e-9)
        # tests/test_represent.py:23-31
        rr, ll = C[k] @ C[k].conj().T, C[k].conj().T @ C[k]
        EL, ER = O.transfer_matrix(AL[k]), O.transfer_matrix(AR[k])
        assert np.abs(EL @ rr.reshape(-1) - rr.reshape(-1)).max() < 1e-10
        assert np.abs(EL.conj().T @ I.reshape(-1) - I.reshape(-1)).max() < 1e-10
        assert np.abs(ER @ I.reshape(-1) - I.reshape(-1)).max() < 1e-9
        assert np.abs(ER.conj().T @ ll.reshape(-1) - ll.reshape(-1)).max() < 1e-9
        ref = O.expectation_values(A[k], _PAULIS)
        assert np.abs(es[k] - ref).max() < 1e-10 and np.abs(es_gen[k] - ref).max() < 1e-10
        assert np.abs(es[k].imag).max() < 1e-12


def test_emu_ansatz_and_energy_match_reference_ground_state_script(emu, golden):
    """The device ansatz evaluator and D = 2 energy routine (compiled for the host) on the gate program of
    scripts/ground_state_finding.py:83-92, against the tensors and energies of the reference's own script
    (tests/golden/ref_ground_state_script.npz, oracle/make_golden_gs.py)."""
    from qmps_b200 import represent as R
    g = golden["ref_ground_state_script"]
    for layers in (1, 2, 4):
        Pn = 4 * layers
        prog = R.GateProgram(2, Pn)
        for l in range(layers):
            prog.rx(0, 4 * l).rx(1, 4 * l + 1).rz(0, 4 * l + 2).rz(1, 4 * l + 3).cnot(0, 1)
        for k in range(6):
            th = np.ascontiguousarray(g[f"p_L{layers}"][k][None, :])
            A = np.zeros((1, 2, 2, 2), complex)
            emu.emu_ansatz_reg2(prog.c_ops(), len(prog), ctypes.c_int64(1), Pn, P(th), -1, ctypes.c_double(0.0), P(A))
            assert np.abs(A[0] - O.unitary_to_tensor(g[f"U_L{layers}"][k])).max() < 1e-13
            for lam in (0.5, 1.0):
                H = np.ascontiguousarray(O.tfim_matrix(lam))
                e = np.zeros(1)
                emu.emu_energy_d2(ctypes.c_int64(1), P(A), P(H), P(e))
                assert abs(e[0] - g[f"eps_L{layers}_lam{lam}"][k]) < 1e-11


def test_gate_programs_match_reference_decompose_methods():
    """Every ansatz class's GateProgram against the gate list its reference class's OWN ``_decompose_`` method
    emits (qmps/represent.py:268-442 executed against a recording cirq stand-in by oracle/make_golden_gates.py):
    same gates in the same order on the same qubits with the same angle / exponent."""
    import json
    from qmps_b200 import represent as R, _lib as L
    names = {L.G_RZ: "rz", L.G_RX: "rx", L.G_RY: "ry", L.G_H: "H", L.G_CNOT: "CNOT", L.G_SWAP: "SWAP", L.G_CZ: "CZ",
             L.G_XPOW: "X**", L.G_ZZPOW: "ZZ**", L.G_XXPOW: "XX**", L.G_YYPOW: "YY**", L.G_X: "X", L.G_Z: "Z"}
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_gate_lists.json")))
    assert len(cases) >= 19
    seen = set()
    for c in cases:
        p = np.array(c["params"])
        gate = getattr(R, c["cls"])(*(([c["D"]] if c["D"] else []) + [p]))
        prog = gate.program()
        assert prog.nq == c["nq"], c["cls"]
        ref_ops = [o for o in c["ops"] if not (o[0] == "SWAP" and o[1][0] == o[1][1])]
        assert len(prog.ops) == len(ref_ops), (c["cls"], c["D"], len(prog.ops), len(ref_ops))
        for (code, q0, q1, param, scale, offset), (name, qubits, value) in zip(prog.ops, ref_ops):
            assert names[code] == name, (c["cls"], c["D"], names[code], name)
            assert [q0, q1][:len(qubits)] == qubits or (len(qubits) == 1 and q0 == qubits[0]), (c["cls"], name, q0, q1, qubits)
            if value is not None:
                angle = scale * p[param] + offset
                assert abs(angle - value) < 1e-15, (c["cls"], name, angle, value)
            else:
                assert param < 0
        seen.add(c["cls"])
    assert len(seen) == 8


def test_emu_fp_d2_edge_cases(emu):
    """Degenerate inputs of the D = 2 register eigen-solver: the zero tensor (eta = 0), a product state (rank-one map,
    eta = 1, eigenvector |0><0|) and NaN input (terminates, flagged QMPS_ST_NO_CONVERGE)."""
    A = np.zeros((3, 2, 2, 2), complex); A[1, 0, 0, 0] = 1.0; A[2] = np.nan
    eta = np.zeros(3, complex); st = np.zeros(3, np.int32); vec = np.zeros((3, 2, 2), complex)
    assert emu.emu_fp_d2(2, ctypes.c_int64(3), P(A), P(A), 0, P(eta), P(st), P(vec)) == 0
    assert eta[0] == 0 and st[0] == 0
    assert abs(eta[1] - 1) < 1e-15 and st[1] == 0 and abs(vec[1, 0, 0] - 1) < 1e-12 and np.abs(vec[1]).sum() < 1 + 1e-9
    assert np.isnan(eta[2].real) and st[2] == 2


def test_c_abi_argument_validation_without_a_device(built):
    """Error behaviour of the C ABI: invalid arguments are rejected with a negative code and a
    message BEFORE any CUDA call, so this runs on the GPU-less container too; empty batches are a no-op success."""
    from qmps_b200 import _lib as L
    lib = L.load()
    ERR_ARG, ERR_UNSUPPORTED = -1, -3
    cases = [
        (lib.qmps_fixed_point(2, 17, 1, 1, 1, 1, 0, 0, None, None, None, None, None, None, L.C128, None), ERR_UNSUPPORTED, "D must be"),
        (lib.qmps_fixed_point(2, 2, 2, 1, 3, 1, 0, 0, None, None, None, None, None, None, L.C128, None), ERR_ARG, "broadcast"),
        (lib.qmps_fixed_point(2, 2, 1, 1, 1, 1, 0, 0, None, None, None, None, None, None, 7, None), ERR_ARG, "dtype"),
        (lib.qmps_tm_power(2, 4, 1, None, None, None, 1, None, L.C128, None), ERR_ARG, "tm_power"),
        (lib.qmps_bw_evolve_cost(4, 3, 1, 1, 4, 1, 1, 1, 1, 1, None, None, None, None, L.C128, None), ERR_ARG, "NK must be 1 or N"),
        (lib.qmps_bw_environment(2, 1, 1, 1, 1, 1, 1, 1, 0, None, None, None, None, L.C128, None), ERR_ARG, "side"),
        (lib.qmps_bw_expectation(1, 1, 1, 1, 3, 1, 1, 1, L.C128, None), ERR_UNSUPPORTED, "2 or 4 qubits"),
        (lib.qmps_cgemm_c64_tc(1, 1, 60, 64, 32, 1, 1, 0, 1, None), ERR_UNSUPPORTED, "multiples of 64"),
        (lib.qmps_set_option(b"no_such_option", 1), ERR_ARG, "unknown option"),
        # round-2 entries
        (lib.qmps_scars_cost(4, 1, 2, 1, 1, 1, None, None, L.C128, None), ERR_ARG, "NC must be 1 or N"),
        (lib.qmps_scars_cost(4, None, 1, None, None, None, None, None, L.C128, None), ERR_ARG, "scars_cost"),
        (lib.qmps_scars_trajectory(None, None, 1, 1, 64, 0.1, 0, 1, None, None, L.C128, None), ERR_ARG, "scars_trajectory"),
        (lib.qmps_env_exact_packed(3, None, 0, None, None), ERR_ARG, "env_exact_packed"),
        (lib.qmps_env_exact_packed_host(3, None, 0, None, 0), ERR_ARG, "env_exact_packed_host"),
        (lib.qmps_get_env_exact_host(3, 1, 1, 1, None, L.C128, 0), ERR_UNSUPPORTED, "unsupported D"),
        (lib.qmps_get_env_exact_host(2, 1, None, None, None, L.C128, 0), ERR_ARG, "get_env_exact_host"),
        (lib.qmps_zgemm_c128_i8(1, 64, 32, 64, None, None, 0, None, None), ERR_ARG, "zgemm_c128_i8"),
        (lib.qmps_tm_apply(2, 4, 1, 1, 1, 8, 8, L.C128, None), ERR_ARG, "must not alias"),
    ]
    for rc, want, msg in cases:
        assert rc == want, (rc, want, msg)
    assert lib.qmps_set_option(b"no_such_option", 1) == ERR_ARG and b"unknown option" in lib.qmps_last_error()
    # empty batches succeed without touching the device
    assert lib.qmps_fixed_point(2, 2, 0, None, 0, None, 0, 0, None, None, None, None, None, None, L.C128, None) == 0
    assert lib.qmps_tm_power(2, 64, 0, None, None, None, 4, None, L.C64, None) == 0
    assert lib.qmps_bw_evolve_cost(0, 1, None, None, 0, None, None, 1, None, None, None, None, None, None, L.C128, None) == 0
    assert lib.qmps_env_exact(2, 2, 0, None, 0, 1, None, None, None, None, L.C128, None) == 0
    assert lib.qmps_scars_cost(0, None, 1, None, None, None, None, None, L.C128, None) == 0
    assert lib.qmps_env_exact_packed(0, None, 0, None, None) == 0
    assert lib.qmps_env_exact_packed_host(0, None, 0, None, 0) == 0
    assert lib.qmps_get_env_exact_host(2, 0, None, None, None, L.C128, 0) == 0


def test_torch_library_ops_registered_cuda_only(built):
    """qmps_b200/ops.py: the custom ops exist, have shape functions (trace as opaque nodes under FakeTensorMode) and
    have NO CPU implementation -- a CPU tensor raises instead of falling back."""
    import torch
    import qmps_b200.ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("env_exact", "fixed_point_cost", "tm_power", "bw_evolve_cost"):
        assert hasattr(torch.ops.qmps_b200, name)
    with pytest.raises(NotImplementedError):
        torch.ops.qmps_b200.env_exact(torch.zeros(1, 2, 2, 2, dtype=torch.complex128))
    with pytest.raises(NotImplementedError):
        torch.ops.qmps_b200.tm_power(torch.zeros(1, 2, 4, 4, dtype=torch.complex64), torch.zeros(1, 2, 4, 4, dtype=torch.complex64), 2)
    with FakeTensorMode():
        A = torch.empty((5, 2, 4, 4), dtype=torch.complex128, device="cuda")
        eta, r, C, st = torch.ops.qmps_b200.env_exact(A)
        assert tuple(r.shape) == (5, 4, 4) and st.dtype == torch.int32 and eta.dtype == torch.complex128
        e2, cost, echo, fid = torch.ops.qmps_b200.fixed_point_cost(A, A[:3], True)
        assert tuple(cost.shape) == (5, 3) and cost.dtype == torch.float64
        V = torch.empty((7, 4, 4), dtype=torch.complex64, device="cuda")
        c = torch.ops.qmps_b200.bw_evolve_cost(V[:1], V[:1], V, V, torch.empty((16, 16), dtype=torch.complex64, device="cuda"))
        assert tuple(c.shape) == (7,) and c.dtype == torch.float32


def test_bench_reference_arm_under_torchrun_prints_one_json_line():
    """`bench.py --impl reference` launched the way the driver launches it for N > 1: rank 0 alone runs and prints
    exactly ONE JSON line on stdout carrying the contract's keys; the other rank exits 0 silently."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29619", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_header_is_plain_c_and_library_binds_from_c(built, tmp_path):
    """include/qmps_b200.h compiles as C99 (and as C++), and a C client can dlopen the library, resolve the
    symbols and get the documented error behaviour -- what a cgo / JNI / FFI stub on the reference side would do."""
    from qmps_b200 import _lib
    hdr = os.path.join(ROOT, "include", "qmps_b200.h")
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        lang = ["-x", "c"] if cc == "gcc" else ["-x", "c++"]
        r = subprocess.run([cc, std, "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", *lang, hdr], capture_output=True, text=True)
        assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    exe = tmp_path / "c_client"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-o", str(exe), os.path.join(ROOT, "tests", "c_abi", "c_client.c"), "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), _lib.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0 and "c_client ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
