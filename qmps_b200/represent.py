"""Ansatz gate lists as data -- the ``qmps.represent`` state-tensor classes without cirq.

The reference classes (qmps/represent.py:188-442) are ``cirq.Gate`` subclasses whose
``_decompose_`` lists gates; ``cirq.unitary(gate)`` then multiplies them out.  Here
each class keeps its constructor signature and gate ORDER, but produces a *gate
program* (rows of ``qmps_gate_op``, include/qmps_b200.h) that the CUDA kernels
interpret for a whole batch of parameter vectors at once.  ``unitary(gate)`` (the
stand-in for ``cirq.unitary``) runs that program on the GPU for one vector.

cirq conventions: qubit 0 is the most significant bit,
rz(t) = exp(-i t Z/2), P**t = exp(i pi t/2) exp(-i pi t P/2).
"""
import numpy as np

from . import _lib as L
from .tools import split_2s, split_3s, split_ns

__all__ = [
    "GateProgram", "Tensor", "StateTensor", "Environment", "FullStateTensor", "FullEnvironment",
    "State", "ShallowQAOAStateTensor", "ShallowCNOTStateTensor", "ShallowCNOTStateTensor_nonuniform",
    "ShallowCNOTStateTensor3", "ExactAfter4", "ShallowFullStateTensor", "StateGate",
    "ShallowEnvironment", "GSFAnsatz", "unitary",
]


class GateProgram:
    """A gate list over ``nq`` qubits reading ``n_params`` parameters.

    ops rows: (code, q0, q1, param, scale, offset); angle = scale*theta[param]+offset.
    """

    def __init__(self, nq, n_params):
        self.nq = int(nq)
        self.n_params = int(n_params)
        self.ops = []

    # -- builders (time order) --
    def _rot(self, code, q, p, scale=1.0, offset=0.0):
        self.ops.append((code, q, q, p, scale, offset))
        return self

    def rz(self, q, p): return self._rot(L.G_RZ, q, p)
    def rx(self, q, p): return self._rot(L.G_RX, q, p)
    def ry(self, q, p): return self._rot(L.G_RY, q, p)
    def h(self, q): return self._rot(L.G_H, q, -1)
    def x(self, q): return self._rot(L.G_X, q, -1)
    def z(self, q): return self._rot(L.G_Z, q, -1)
    def xpow(self, q, p): return self._rot(L.G_XPOW, q, p)

    def _two(self, code, qa, qb, p=-1):
        self.ops.append((code, qa, qb, p, 1.0, 0.0))
        return self

    def cnot(self, c, t): return self._two(L.G_CNOT, c, t)
    def swap(self, a, b): return self if a == b else self._two(L.G_SWAP, a, b)
    def cz(self, a, b): return self._two(L.G_CZ, a, b)
    def zzpow(self, a, b, p): return self._two(L.G_ZZPOW, a, b, p)
    def xxpow(self, a, b, p): return self._two(L.G_XXPOW, a, b, p)
    def yypow(self, a, b, p): return self._two(L.G_YYPOW, a, b, p)

    @property
    def bond_dim(self):
        return 2 ** (self.nq - 1)

    def c_ops(self):
        return L.make_ops(self.ops)

    def __len__(self):
        return len(self.ops)


def _nq(bond_dim):
    return int(np.log2(bond_dim)) + 1


class _ProgramGate:
    """Common behaviour of the parameterised state tensors."""

    def __init__(self, bond_dim, params, symbol="U"):
        self.βγs = params
        self.params = np.asarray(params, dtype=np.float64)
        self.p = len(params)
        self.n_qubits = _nq(bond_dim)
        self.D = bond_dim
        self.symbol = symbol

    def num_qubits(self):
        return self.n_qubits

    def program(self):
        raise NotImplementedError

    def _unitary_(self):
        from .batched import ansatz_unitaries_host
        return ansatz_unitaries_host(self.program(), self.params[None, :])[0]

    def _circuit_diagram_info_(self, args=None):
        return [self.symbol] * self.n_qubits

    def _cnot_ladder_reversed(self, prog):
        for i in reversed(range(self.n_qubits - 1)):
            prog.cnot(i, i + 1)


class ShallowFullStateTensor(_ProgramGate):
    """15-parameter two-qubit gate (qmps/represent.py:382-404)."""

    def __init__(self, bond_dim, βγs, symbol="U"):
        super().__init__(bond_dim, βγs, symbol)

    def program(self):
        g = GateProgram(self.n_qubits, 15)
        g.rz(0, 0).rx(0, 1).rz(0, 2).rz(1, 3).rx(1, 4).rz(1, 5)
        g.cnot(0, 1).ry(0, 6).cnot(1, 0).ry(0, 7).rz(1, 8).cnot(0, 1)
        g.rz(0, 9).rx(0, 10).rz(0, 11).rz(1, 12).rx(1, 13).rz(1, 14)
        return g


class ShallowCNOTStateTensor(_ProgramGate):
    """qmps/represent.py:288-310: (beta, gamma) per layer."""

    def __init__(self, bond_dim, βγs):
        super().__init__(bond_dim, βγs)

    @staticmethod
    def params_per_iter():
        return 2

    def program(self):
        n = self.n_qubits
        g = GateProgram(n, self.p)
        for k in range(0, self.p - self.p % 2, 2):
            for q in range(n): g.rz(q, k)
            for q in range(n): g.rx(q, k + 1)
            g.h(0)
            self._cnot_ladder_reversed(g)
        return g


class ShallowCNOTStateTensor_nonuniform(_ProgramGate):
    """qmps/represent.py:312-332: 2*n_qubits parameters per layer."""

    def __init__(self, bond_dim, βγs):
        super().__init__(bond_dim, βγs)

    @staticmethod
    def params_per_iter(D):
        return int((np.log2(D) + 1) * 2)

    def program(self):
        n = self.n_qubits
        g = GateProgram(n, self.p)
        for k in range(0, self.p - self.p % (2 * n), 2 * n):
            for q in range(n): g.rz(q, k + q)
            for q in range(n): g.rx(q, k + n + q)
            self._cnot_ladder_reversed(g)
        return g


class ShallowCNOTStateTensor3(_ProgramGate):
    """qmps/represent.py:334-354."""

    def __init__(self, bond_dim, βγs):
        super().__init__(bond_dim, βγs)

    def program(self):
        n = self.n_qubits
        g = GateProgram(n, self.p)
        for k in range(0, self.p - self.p % 3, 3):
            for q in range(n): g.rz(q, k)
            for q in range(n): g.rx(q, k + 1)
            for q in range(n): g.rz(q, k + 2)
            g.h(0)
            self._cnot_ladder_reversed(g)
        return g


class ShallowQAOAStateTensor(_ProgramGate):
    """qmps/represent.py:268-285: X**beta on every qubit, ZZ**gamma on neighbours."""

    def __init__(self, bond_dim, βγs):
        super().__init__(bond_dim, βγs)

    def program(self):
        n = self.n_qubits
        g = GateProgram(n, self.p)
        for k in range(0, self.p - self.p % 2, 2):
            for q in range(n): g.xpow(q, k)
            for q in range(n - 1): g.zzpow(q, q + 1, k + 1)
        return g


class ShallowEnvironment(ShallowQAOAStateTensor):
    """qmps/represent.py:425-442: the same QAOA gate list on 2 log2(D) qubits."""

    def __init__(self, bond_dim, βγs):
        super().__init__(bond_dim, βγs)
        self.n_qubits = 2 * int(np.log2(bond_dim))
        self.symbol = "V"


class ExactAfter4(_ProgramGate):
    """qmps/represent.py:356-380."""

    def __init__(self, bond_dim, βγs):
        super().__init__(bond_dim, βγs)

    @staticmethod
    def params_per_iter():
        return 6

    def program(self):
        n = self.n_qubits
        g = GateProgram(n, self.p)
        for k in range(0, self.p - self.p % 6, 6):
            a, b, c, d, e, f = range(k, k + 6)
            g.rz(0, a).rz(1, d).rx(0, b).rx(1, e).rz(0, c).rz(1, f)
            self._cnot_ladder_reversed(g)
            for i in range(n):
                g.swap(i, i + 1 if i != n - 1 else 0)
        return g


class StateGate(_ProgramGate):
    """qmps/represent.py:406-423: six-parameter two-qubit gate."""

    def __init__(self, βγs, symbol="U"):
        super().__init__(2, βγs, symbol)

    def program(self):
        g = GateProgram(2, max(self.p, 6))
        g.rx(0, 0).rx(1, 1).rz(0, 2).rz(1, 3).xxpow(0, 1, 4).yypow(0, 1, 5)
        return g


class GSFAnsatz(_ProgramGate):
    """The cirq-free D=2 ansatz of scripts/ground_state_finding.py:83-92: per four
    parameters Rx(w) x Rx(x), Rz(u) x Rz(v), CNOT (zero-padded to a multiple of 4)."""

    def __init__(self, params):
        params = list(params)
        if len(params) % 4:
            params = params + [0.0] * (4 - len(params) % 4)
        super().__init__(2, params)

    def program(self):
        g = GateProgram(2, self.p)
        for k in range(0, self.p, 4):
            g.rx(0, k).rx(1, k + 1).rz(0, k + 2).rz(1, k + 3).cnot(0, 1)
        return g


# ---- fixed (non-parameterised) tensors: plain data holders ------------------------
class Tensor:
    """qmps/represent.py:188-207 without the cirq base class."""

    def __init__(self, unitary, symbol):
        self.U = np.asarray(unitary)
        self.n_qubits = int(np.log2(self.U.shape[0]))
        self.symbol = symbol

    def _unitary_(self):
        return self.U

    def num_qubits(self):
        return self.n_qubits

    def _circuit_diagram_info_(self, args=None):
        return [self.symbol] * self.n_qubits

    def __pow__(self, power, modulo=None):
        if power == -1:
            return self.__class__(self.U.conj().T, symbol=self.symbol + "†")
        return self.__class__(np.linalg.matrix_power(self.U, power), symbol=self.symbol)


class StateTensor(Tensor):
    pass


class Environment(Tensor):
    pass


class FullStateTensor(StateTensor):
    def __init__(self, unitary, symbol="U"):
        super().__init__(unitary, symbol)


class FullEnvironment(Environment):
    def __init__(self, unitary, symbol="V"):
        super().__init__(unitary, symbol)


class State:
    """qmps/represent.py:251-265: holds (u, v, n); the circuit it stands for is
    v on qubits [n, n+v), then u on [i, i+u) for i = n-1..0."""

    def __init__(self, u, v, n=1):
        self.u, self.v, self.n_phys_qubits = u, v, n
        self.bond_dim = int(2 ** (u.num_qubits() - 1))

    def num_qubits(self):
        return self.n_phys_qubits + self.v.num_qubits()


def unitary(gate):
    """Stand-in for ``cirq.unitary(gate)`` on the classes of this module."""
    return gate._unitary_()
