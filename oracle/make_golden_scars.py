"""Generate tests/golden/ref_scars.npz: the reference's own PXP scar cost functions
(``scars.py:76-111`` ``scars_time_evolve_cost_function`` -- circuit-parameterised gates -- and ``:113-155``
``scars_cost_fun_alternate`` -- ``tensor_to_unitary`` gates), its ansatz tensor ``A``, Hamiltonian ``H``, evolution
gate ``W`` and the classical TDVP right-hand side ``func_list``, cut out of the reference source with ``ast`` and
executed unmodified on seeded inputs.

Stated here, not the reference's: the minimal state-vector stand-in for cirq of ``make_golden_obj.py`` plus the four
gates this file needs (``ZPowGate``, ``CNotPowGate``, ``S``, ``X``; cirq's conventions), and xmps'
``Map(...).right_fixed_point()`` supplied by the oracle (zgeev gauge, unit norm -- what the notebooks record).

Run in the build container only:  ``python oracle/make_golden_scars.py``.
"""
import ast
import os
import sys

import numpy as np
from scipy.linalg import expm, null_space

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden_obj as MO           # noqa: E402


def cut_any(path, names, namespace):
    """functions, classes AND module-level assignments (the reference defines A, H, W as lambdas)"""
    import unicodedata
    names = [unicodedata.normalize("NFKC", n) for n in names]      # Python normalises identifiers (U+03D5 -> U+03C6)
    src = open(os.path.join(REF, path)).read()
    lines = src.splitlines()
    for node in ast.parse(src).body:
        name = None
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            name = node.name
        elif isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
        if name in names:
            exec(compile("\n" * (node.lineno - 1) + "\n".join(lines[node.lineno - 1:node.end_lineno]), path, "exec"), namespace)
    assert all(n in namespace for n in names), [n for n in names if n not in namespace]
    return namespace


def main():
    import make_golden as MG
    import oracle as O
    ref_tools = MG.load_reference_tools()
    cirq = MO.mini_cirq()
    cirq.X = MO.Prim([[0, 1], [1, 0]])
    cirq.S = MO.Prim(np.diag([1, 1j]))
    cirq.ZPowGate = lambda exponent=1.0: MO.Prim(np.diag([1, np.exp(1j * np.pi * exponent)]))

    def cnotpow(exponent=1.0):
        t = exponent
        c, s = np.cos(np.pi * t / 2), np.sin(np.pi * t / 2)
        xt = np.exp(1j * np.pi * t / 2) * np.array([[c, -1j * s], [-1j * s, c]])
        return MO.Prim(np.block([[np.eye(2), np.zeros((2, 2))], [np.zeros((2, 2)), xt]]))
    cirq.CNotPowGate = cnotpow
    cirq.inverse = lambda op: MO.Op(MO.Prim(MO.unitary(op.gate).conj().T), op.qubits) if isinstance(op, MO.Op) else op ** -1

    class Map:
        def __init__(self, A, B):
            self.A, self.B = A, B

        def right_fixed_point(self):
            return O.right_fixed_point(self.A, self.B)

    ns = dict(np=np, cirq=cirq, log2=np.log2, null_space=null_space, expm=expm, Map=Map, kron=np.kron,
              tensor_to_unitary=ref_tools.tensor_to_unitary)
    import functools
    ns["reduce"] = functools.reduce
    MO.cut("qmps/represent.py", ["Tensor", "Environment"], ns)
    MO.cut("qmps/time_evolve_tools.py", ["merge", "put_env_on_left_site", "put_env_on_right_site"], ns)
    for k in ("sin", "cos", "tan", "arcsin", "pi"):
        ns[k] = getattr(np, k)
    cut_any("scars.py", ["multi_tensor", "P", "X", "n", "I", "H", "W", "ScarsAnsatz", "ScarGate", "A",
                         "scars_time_evolve_cost_function", "scars_cost_fun_alternate", "dθdt", "dϕdt", "func_list"], ns)
    rng = np.random.default_rng(73)
    mus, dts = np.array([0.325, 0.0, 1.0]), np.array([0.2, 0.05, 0.4])
    cur = rng.normal(size=(3, 4))
    par = np.concatenate([cur[:, None, :] + 0.1 * rng.normal(size=(3, 3, 4)), rng.normal(size=(3, 1, 4))], axis=1)   # 3 near, 1 far
    Wm = np.stack([MO.unitary(ns["W"](m, t)) for m, t in zip(mus, dts)])
    Hm = np.stack([ns["H"](m) for m in mus])
    c_circ = np.zeros((3, 4, 3)); c_alt = np.zeros((3, 4, 3))
    for a in range(3):
        for b in range(4):
            for w in range(3):
                ham = ns["W"](mus[w], dts[w])
                c_circ[a, b, w] = float(ns["scars_time_evolve_cost_function"](par[a, b], cur[a], ham))
                c_alt[a, b, w] = float(ns["scars_cost_fun_alternate"](par[a, b], cur[a], ham))
    tens = np.stack([ns["A"](t, p) for t, p in cur[:, :2]])
    rhs = np.stack([np.array(ns["func_list"](list(c), 0.0, 0.325), dtype=float) for c in cur])
    gates = np.stack([MO.unitary(ns["ScarGate"](list(c))) for c in cur])
    np.savez_compressed(os.path.join(OUT, "ref_scars.npz"), mus=mus, dts=dts, cur=cur, par=par, W=Wm, H=Hm,
                        cost_circuit=c_circ, cost_alternate=c_alt, A=tens, rhs=rhs, scar_gate=gates)
    print("wrote ref_scars.npz; max |circuit - alternate| =", np.abs(c_circ - c_alt).max(), "sample", c_circ[0, :, 0])


if __name__ == "__main__":
    main()
