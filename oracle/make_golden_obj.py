"""Generate tests/golden/ref_loschmidt_obj.npz: the reference's own Loschmidt / TDVP-step cost
``obj(p, A, WW)`` (``qmps/loschmidts/time_evo.py:75-116`` = ``scripts/loschmidt.py:209-239``), cut out with
``ast`` and executed unmodified, together with the helpers it calls from the same file (``merge``,
``put_env_on_left_site``, ``put_env_on_right_site``, ``gate``), ``Tensor`` / ``Environment`` /
``ShallowFullStateTensor`` from ``qmps/represent.py`` and ``unitary_to_tensor`` / ``tensor_to_unitary`` from the
reference's ``qmps/tools.py``; and ``get_overlap_exact(p1, p2)`` of ``qmps/time_evolve_tools.py:84-91`` likewise.

What is NOT the reference's and is therefore stated here: a minimal state-vector stand-in for cirq (gate matrices
in cirq's conventions, big-endian LineQubits, ``Circuit``, ``Simulator.simulate(...).final_state``, ``unitary``,
``inverse``) and xmps' ``iMPS(...).left_canonicalise()`` / ``Map(...).right_fixed_point()`` supplied by the oracle.
What it pins is the reference's circuit layout -- which unitaries sit on which of the six qubits, in which order,
the environment embeddings and the ``sqrt(2|amplitude|)`` read-out -- and hence the identity
``obj = -sqrt|eta_2|`` that the CUDA path computes.

Run in the build container only:  ``python oracle/make_golden_obj.py``.
"""
import ast
import os
import sys
import types

import numpy as np
from scipy.linalg import expm, null_space

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)


# ------------------------------------------------------------------ a minimal cirq (stated conventions)
class Op:
    def __init__(self, gate, qubits):
        self.gate, self.qubits = gate, tuple(qubits)


class Gate:
    def __call__(self, *qubits):
        return Op(self, qubits)

    def on(self, *qubits):
        return Op(self, qubits)


class Prim(Gate):
    def __init__(self, mat):
        self.mat = np.asarray(mat, dtype=complex)

    def _unitary_(self):
        return self.mat

    def num_qubits(self):
        return int(np.log2(self.mat.shape[0]))


def _flatten(x):
    if isinstance(x, Op):
        return [x]
    out = []
    for y in x:
        out += _flatten(y)
    return out


def _apply(state, mat, qubits, n):
    k = len(qubits)
    psi = state.reshape((2,) * n)
    psi = np.moveaxis(psi, qubits, range(k)).reshape(2 ** k, -1)
    psi = (mat @ psi).reshape((2,) * n)
    return np.moveaxis(psi, range(k), qubits).reshape(-1)


def unitary(g):
    if hasattr(g, "_unitary_"):
        return np.asarray(g._unitary_(), dtype=complex)
    n = g.num_qubits()
    U = np.eye(2 ** n, dtype=complex)
    for op in _flatten(g._decompose_(list(range(n)))):
        m = unitary(op.gate)
        U = np.stack([_apply(U[:, c], m, list(op.qubits), n) for c in range(2 ** n)], axis=1)
    return U


def mini_cirq():
    c = types.SimpleNamespace()
    c.Gate = Gate
    rot = lambda P: (lambda th: Prim(expm(-0.5j * th * P)))             # noqa: E731
    c.rz, c.rx, c.ry = rot(Z), rot(X), rot(Y)
    c.H = Prim(np.array([[1, 1], [1, -1]]) / np.sqrt(2))
    c.CNOT = Prim([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    c.SWAP = Prim([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    c.unitary = unitary
    c.inverse = lambda g: g ** -1
    c.LineQubit = types.SimpleNamespace(range=lambda *a: list(range(*a)))

    class Circuit:
        def __init__(self, ops=()):
            self.ops = _flatten(ops)

        @classmethod
        def from_ops(cls, *ops):
            return cls(list(ops))
    c.Circuit = Circuit

    class Simulator:
        def __init__(self, dtype=None):
            pass

        def simulate(self, C):
            n = 1 + max(q for op in C.ops for q in op.qubits)
            psi = np.zeros(2 ** n, dtype=complex)
            psi[0] = 1
            for op in C.ops:
                psi = _apply(psi, unitary(op.gate), list(op.qubits), n)
            return types.SimpleNamespace(final_state=psi)
    c.Simulator = Simulator
    return c


def cut(path, names, namespace):
    src = open(os.path.join(REF, path)).read()
    lines = src.splitlines()
    for node in ast.parse(src).body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile("\n".join(lines[node.lineno - 1:node.end_lineno]), f"{path}:{node.lineno}", "exec"), namespace)
    assert all(n in namespace for n in names), [n for n in names if n not in namespace]
    return namespace


def main():
    import make_golden as MG
    import oracle as O
    ref_tools = MG.load_reference_tools()

    class iMPS:                                   # xmps stand-ins (oracle restatements)
        def __init__(self, data):
            self.data = data

        def left_canonicalise(self):
            return [O.left_canonicalise(self.data[0])]

    class Map:
        def __init__(self, A, B):
            self.A, self.B = A, B

        def right_fixed_point(self):
            return O.right_fixed_point(self.A, self.B)

        def left_fixed_point(self):
            return O.left_fixed_point(self.A, self.B)

    cirq = mini_cirq()
    ns = dict(np=np, cirq=cirq, log2=np.log2, null_space=null_space, iMPS=iMPS, Map=Map,
              unitary_to_tensor=ref_tools.unitary_to_tensor, tensor_to_unitary=ref_tools.tensor_to_unitary)
    cut("qmps/represent.py", ["Tensor", "Environment", "ShallowFullStateTensor"], ns)
    cut("qmps/loschmidts/time_evo.py", ["merge", "put_env_on_left_site", "put_env_on_right_site", "gate", "obj"], ns)
    rng = np.random.default_rng(41)
    H = O.tfim_matrix(0.2)
    Ws = np.stack([np.eye(4, dtype=complex), expm(-1j * H * 2 * 0.02), expm(-1j * H * 2 * 0.3)])
    p0 = rng.normal(size=(3, 15))
    ps = rng.normal(size=(4, 15))
    ps[0] = p0[0]                                  # the state itself: cost -1 at W = 1
    ps[1] = p0[0] + 0.05 * rng.normal(size=15)
    A0 = np.stack([O.left_canonicalise(ref_tools.unitary_to_tensor(unitary(ns["gate"](p))))  for p in p0])
    vals = np.zeros((3, 4, 3))
    for a in range(3):
        for b in range(4):
            for w in range(3):
                vals[a, b, w] = float(np.real(ns["obj"](ps[b], A0[a], Ws[w])))
    # get_overlap_exact(p1, p2) of qmps/time_evolve_tools.py:84-91, same stand-ins (its `gate` default is the
    # ShallowFullStateTensor factory of qmps/rotosolve.py:14-17, identical to time_evo.py's)
    ns2 = dict(ns)
    cut("qmps/time_evolve_tools.py", ["get_overlap_exact"], ns2)
    ov = np.zeros((3, 4))
    ov_r = np.zeros((3, 4, 2, 2), dtype=complex)
    for a in range(3):
        for b in range(4):
            f, r = ns2["get_overlap_exact"](p0[a], ps[b])
            ov[a, b], ov_r[a, b] = f, r
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ref_loschmidt_obj.npz"), p0=p0, ps=ps, A0=A0, Ws=Ws, obj=vals,
                        overlap=ov, overlap_r=ov_r,
                        U_gate=np.stack([unitary(ns["gate"](p)) for p in ps]))
    print("wrote ref_loschmidt_obj.npz", vals.shape, "obj(state itself, W=1) =", vals[0, 0, 0])


if __name__ == "__main__":
    main()
