"""Oracle, part 2: the cirq gate conventions and the in-repo ansaetze theta -> U.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  cirq is an un-vendored,
un-pinned dependency of the reference and is absent here, so its published gate
definitions are restated: big-endian qubit order (qubit 0 is the
most significant bit, as in ``np.kron``), ``rz(t) = exp(-i t Z/2)`` (likewise
rx, ry), ``P**t = exp(i pi t/2) exp(-i pi t P/2)`` for P in {X, ZZ, XX, YY}.
In-repo matrix restatements with the same conventions:
``scripts/ground_state_finding.py:74-92`` (Rx/Ry/Rz/ansatz).
[cirq: parity unpinned]
"""
import numpy as np

__all__ = [
    "I2", "PX", "PY", "PZ", "rx", "ry", "rz", "hadamard", "cnot", "swap",
    "pauli_pow", "xpow", "zzpow", "xxpow", "yypow", "on_qubits",
    "circuit_unitary", "simulate", "split_ns",
    "shallow_full_state_tensor", "shallow_cnot_state_tensor",
    "shallow_cnot_state_tensor_nonuniform", "shallow_cnot_state_tensor3",
    "shallow_qaoa_state_tensor", "exact_after4", "state_gate", "gsf_ansatz",
    "state_circuit",
]

I2 = np.eye(2, dtype=np.complex128)
PX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
PY = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
PZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)


def _rot(P, t):
    return np.cos(t / 2) * np.eye(P.shape[0]) - 1j * np.sin(t / 2) * P


def rx(t):
    return _rot(PX, t)


def ry(t):
    return _rot(PY, t)


def rz(t):
    return _rot(PZ, t)


def hadamard():
    return np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2)


def cnot():
    """control = first (more significant) qubit."""
    return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)


def swap():
    return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)


def pauli_pow(P, t):
    """cirq's EigenGate power for an involution P: eigenvalue +1 -> 1, -1 -> e^{i pi t}."""
    return np.exp(1j * np.pi * t / 2) * _rot(P, np.pi * t)


def xpow(t):
    return pauli_pow(PX, t)


def zzpow(t):
    return pauli_pow(np.kron(PZ, PZ), t)


def xxpow(t):
    return pauli_pow(np.kron(PX, PX), t)


def yypow(t):
    return pauli_pow(np.kron(PY, PY), t)


def _apply(psi, g, qubits, n):
    """Apply the k-qubit matrix g to the listed qubits of an n-qubit tensor whose
    first n axes are the qubits (extra trailing axes are carried along)."""
    k = len(qubits)
    g = np.asarray(g, dtype=np.complex128).reshape((2,) * (2 * k))
    out = np.tensordot(g, psi, axes=(list(range(k, 2 * k)), list(qubits)))
    # tensordot puts the k output axes first; move them back to `qubits`
    return np.moveaxis(out, list(range(k)), list(qubits))


def on_qubits(g, qubits, n):
    """Dense 2^n x 2^n matrix of gate g acting on `qubits` (big-endian)."""
    dim = 2 ** n
    psi = np.eye(dim, dtype=np.complex128).reshape((2,) * n + (dim,))
    return _apply(psi, g, qubits, n).reshape(dim, dim)


def circuit_unitary(ops, n):
    """ops: iterable of (matrix, qubits) in time order."""
    dim = 2 ** n
    psi = np.eye(dim, dtype=np.complex128).reshape((2,) * n + (dim,))
    for g, qs in ops:
        psi = _apply(psi, g, qs, n)
    return psi.reshape(dim, dim)


def simulate(ops, n):
    """Final state of the circuit applied to |0...0> (cirq.Simulator().simulate(C).final_state)."""
    psi = np.zeros(2 ** n, dtype=np.complex128)
    psi[0] = 1
    psi = psi.reshape((2,) * n)
    for g, qs in ops:
        psi = _apply(psi, g, qs, n)
    return psi.reshape(-1)


def split_ns(x, n):
    """qmps/tools.py:167-170."""
    return [x[i:i + n] for i in range(0, len(x), n)]


def _nq(D):
    return int(round(np.log2(D))) + 1


def _cnot_ladder_reversed(n):
    """list(reversed([CNOT(q[i], q[i+1]) for i in range(n-1)])) (represent.py:304)."""
    return [(cnot(), (i, i + 1)) for i in reversed(range(n - 1))]


# ---- ansaetze (qmps/represent.py:268-423) ---------------------------------
def shallow_full_state_tensor(p):
    """15-parameter two-qubit gate (qmps/represent.py:392-401)."""
    ops = [(rz(p[0]), (0,)), (rx(p[1]), (0,)), (rz(p[2]), (0,)),
           (rz(p[3]), (1,)), (rx(p[4]), (1,)), (rz(p[5]), (1,)),
           (cnot(), (0, 1)),
           (ry(p[6]), (0,)),
           (cnot(), (1, 0)),
           (ry(p[7]), (0,)), (rz(p[8]), (1,)),
           (cnot(), (0, 1)),
           (rz(p[9]), (0,)), (rx(p[10]), (0,)), (rz(p[11]), (0,)),
           (rz(p[12]), (1,)), (rx(p[13]), (1,)), (rz(p[14]), (1,))]
    return circuit_unitary(ops, 2)


def shallow_cnot_state_tensor(D, p):
    """qmps/represent.py:300-307: per (beta, gamma): rz(beta) on all, rx(gamma) on
    all, H on qubit 0, reversed CNOT ladder."""
    n = _nq(D)
    ops = []
    for b, g in split_ns(list(p), 2):
        ops += [(rz(b), (q,)) for q in range(n)]
        ops += [(rx(g), (q,)) for q in range(n)]
        ops += [(hadamard(), (0,))]
        ops += _cnot_ladder_reversed(n)
    return circuit_unitary(ops, n)


def shallow_cnot_state_tensor_nonuniform(D, p):
    """qmps/represent.py:325-329: per layer of 2n parameters: rz(p[i]) on qubit i,
    rx(p[n+i]) on qubit i, reversed CNOT ladder."""
    n = _nq(D)
    ops = []
    for layer in split_ns(list(p), 2 * n):
        ops += [(rz(layer[q]), (q,)) for q in range(n)]
        ops += [(rx(layer[n + q]), (q,)) for q in range(n)]
        ops += _cnot_ladder_reversed(n)
    return circuit_unitary(ops, n)


def shallow_cnot_state_tensor3(D, p):
    """qmps/represent.py:344-352."""
    n = _nq(D)
    ops = []
    for b, g, w in split_ns(list(p), 3):
        ops += [(rz(b), (q,)) for q in range(n)]
        ops += [(rx(g), (q,)) for q in range(n)]
        ops += [(rz(w), (q,)) for q in range(n)]
        ops += [(hadamard(), (0,))]
        ops += _cnot_ladder_reversed(n)
    return circuit_unitary(ops, n)


def shallow_qaoa_state_tensor(D, p):
    """qmps/represent.py:279-282: X**beta on every qubit, ZZ**gamma on neighbours."""
    n = _nq(D)
    ops = []
    for b, g in split_ns(list(p), 2):
        ops += [(xpow(b), (q,)) for q in range(n)]
        ops += [(zzpow(g), (q, q + 1)) for q in range(n - 1)]
    return circuit_unitary(ops, n)


def exact_after4(D, p):
    """qmps/represent.py:370-377 (6 parameters per layer on qubits 0,1; reversed
    CNOT ladder; cyclic SWAP chain)."""
    n = _nq(D)
    ops = []
    for a, b, c, d, e, f in split_ns(list(p), 6):
        ops += [(rz(a), (0,)), (rz(d), (1,)), (rx(b), (0,)), (rx(e), (1,)),
                (rz(c), (0,)), (rz(f), (1,))]
        ops += _cnot_ladder_reversed(n)
        ops += [(swap(), (i, i + 1 if i != n - 1 else 0)) for i in range(n)]
    return circuit_unitary(ops, n)


def state_gate(p):
    """qmps/represent.py:416-420."""
    a, b, c, d, e, f = p[:6]
    ops = [(rx(a), (0,)), (rx(b), (1,)), (rz(c), (0,)), (rz(d), (1,)),
           (xxpow(e), (0, 1)), (yypow(f), (0, 1))]
    return circuit_unitary(ops, 2)


def gsf_ansatz(p):
    """scripts/ground_state_finding.py:83-92: per 4 parameters (w,x,u,v):
    Rx(w) (x) Rx(x), then Rz(u) (x) Rz(v), then CNOT; zero-padded to a multiple of 4."""
    p = list(p)
    if len(p) % 4:
        p = p + [0.0] * (4 - len(p) % 4)
    ops = []
    for w, x, u, v in split_ns(p, 4):
        ops += [(rx(w), (0,)), (rx(x), (1,)), (rz(u), (0,)), (rz(v), (1,)), (cnot(), (0, 1))]
    return circuit_unitary(ops, 2)


def state_circuit(U, V, n_phys=2):
    """``State(U, V, n)`` (qmps/represent.py:258-262): V on qubits [n, n+v), then U
    on [i, i+u) for i = n-1 ... 0.  Returns (ops, total_qubits)."""
    u = int(round(np.log2(U.shape[0])))
    v = int(round(np.log2(V.shape[0])))
    ops = [(V, tuple(range(n_phys, n_phys + v)))]
    ops += [(U, tuple(range(i, i + u))) for i in reversed(range(n_phys))]
    return ops, n_phys + v
