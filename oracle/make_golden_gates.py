"""Generate tests/golden/ref_gate_lists.json: the gate lists of the reference's ansatz classes, recorded by
running the reference's OWN ``_decompose_`` methods (``qmps/represent.py:268-442``) against a recording stand-in
for cirq.

The class definitions are cut out of ``qmps/represent.py`` with ``ast`` and executed unmodified; ``split_2s`` /
``split_3s`` / ``split_ns`` likewise from ``qmps/tools.py:161-174``.  The stand-in ``cirq`` implements only what
those methods touch -- ``cirq.Gate``, ``rz/rx/ry``, ``X``, ``H``, ``CNOT``, ``SWAP``, ``ZZ``, ``XX``, ``YY`` and
``**`` -- and records ``(name, qubits, value)`` instead of building matrices.  What this pins is the reference's
gate ORDER, qubit assignment and parameter mapping; the gate MATRICES (cirq's conventions:
``rz = exp(-i Z theta / 2)``, ``X**t``, ``XX**t`` ...) remain stated in oracle/gates.py.

Run in the build container only:  ``python oracle/make_golden_gates.py``.
"""
import ast
import json
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


class Op:
    def __init__(self, name, qubits, value=None):
        self.name, self.qubits, self.value = name, tuple(qubits), value

    def __pow__(self, e):                       # cirq.X(q) ** beta, cirq.ZZ(a, b) ** gamma
        return Op(self.name + "**", self.qubits, float(e))


class G:
    """A gate object: callable on qubits, optionally raised to a power first ((cirq.XX ** e)(*qubits))."""
    def __init__(self, name, value=None):
        self.name, self.value = name, value

    def __call__(self, *qubits):
        return Op(self.name, qubits, self.value)

    def __pow__(self, e):
        return G(self.name + "**", float(e))


def recording_cirq():
    c = types.SimpleNamespace()
    c.Gate = object
    for r in ("rz", "rx", "ry"):
        setattr(c, r, (lambda name: (lambda theta: G(name, float(theta))))(r))
    for g in ("X", "Z", "H", "CNOT", "SWAP", "CZ", "ZZ", "XX", "YY"):
        setattr(c, g, G(g))
    return c


def cut(path, names, namespace):
    src = open(os.path.join(REF, path)).read()
    lines = src.splitlines()
    for node in ast.parse(src).body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile("\n".join(lines[node.lineno - 1:node.end_lineno]), f"{path}:{node.lineno}", "exec"), namespace)
    assert all(n in namespace for n in names)
    return namespace


def flatten(x):
    if isinstance(x, Op):
        return [x]
    out = []
    for y in x:
        out += flatten(y)
    return out


CLASSES = ["ShallowQAOAStateTensor", "ShallowCNOTStateTensor", "ShallowCNOTStateTensor_nonuniform",
           "ShallowCNOTStateTensor3", "ExactAfter4", "ShallowFullStateTensor", "StateGate", "ShallowEnvironment"]


def main():
    ns = dict(cirq=recording_cirq(), log2=np.log2, np=np)
    cut("qmps/tools.py", ["split_2s", "split_3s", "split_ns"], ns)
    cut("qmps/represent.py", CLASSES, ns)
    rng = np.random.default_rng(31)
    cases = []

    def record(cls, D, nparams, *ctor_extra):
        p = rng.normal(size=nparams)
        gate = ns[cls](*(([D] if D else []) + [p]))
        nq = gate.num_qubits()
        ops = flatten(gate._decompose_(list(range(nq))))
        cases.append({"cls": cls, "D": D, "params": p.tolist(), "nq": nq,
                      "ops": [[o.name, list(o.qubits), o.value] for o in ops]})
    for D in (2, 4, 8):
        nq = int(np.log2(D)) + 1
        record("ShallowQAOAStateTensor", D, 6)
        record("ShallowCNOTStateTensor", D, 6)
        record("ShallowCNOTStateTensor_nonuniform", D, 2 * nq * 3)
        record("ShallowCNOTStateTensor3", D, 9)
        record("ShallowEnvironment", D, 4)
    record("ExactAfter4", 2, 12)
    record("ExactAfter4", 4, 12)
    record("ShallowFullStateTensor", 2, 15)
    record("StateGate", 0, 6)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "ref_gate_lists.json"), "w") as f:
        json.dump(cases, f)
    print("wrote ref_gate_lists.json:", len(cases), "cases,", sum(len(c["ops"]) for c in cases), "ops")


if __name__ == "__main__":
    main()
