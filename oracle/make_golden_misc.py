"""Generate tests/golden/ref_misc.npz from individual pure functions of reference modules that cannot be
imported whole (cirq.Gate subclasses and xmps imports at module top).  The named definitions are cut out of the
reference source with ``ast`` and executed UNMODIFIED in a namespace holding numpy / scipy and the few names
they read:

* ``qmps/time_evolve_tools.py``: ``merge`` (:20-23), ``put_env_on_left_site`` (:38-53), ``get_env_off_left_site``
  (:55-57), ``put_env_on_right_site`` (:59-70), ``get_env_off_right_site`` (:72-74)
* ``qmps/ground_state.py``: class ``Hamiltonian`` (:66-88; ``to_matrix`` and the single-letter key expansion)
* ``qmps/rotosolve.py``: ``rotosolve`` (:154-181) and ``double_rotosolve`` (:183-241) on a deterministic
  state function

Stated, not taken from the reference: the Pauli matrices ``xmps.spin.paulis(0.5)`` would return, and
``cirq.unitary(cirq.SWAP)`` (the 4x4 swap).  Plot calls are absorbed by a permissive stub.

Run in the build container only:  ``python oracle/make_golden_misc.py``.
"""
import ast
import os
import sys
import types
from functools import reduce
from itertools import product

import numpy as np
from scipy.linalg import expm, null_space
from scipy.optimize import minimize_scalar

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)


class _Anything:
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def cut(path, names, namespace):
    """exec the top-level definitions `names` of the reference file `path`, unmodified, in `namespace`."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    lines = src.splitlines()
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            code = "\n".join(lines[node.lineno - 1:node.end_lineno])
            exec(compile(code, f"{path}:{node.lineno}-{node.end_lineno}", "exec"), namespace)
    missing = [n for n in names if n not in namespace]
    assert not missing, missing
    return namespace


def main():
    rng = np.random.default_rng(21)
    out = {}
    # ---- time_evolve_tools
    cirq = types.SimpleNamespace(SWAP="SWAP", unitary=lambda g: SWAP)
    tet = cut("qmps/time_evolve_tools.py",
              ["merge", "put_env_on_left_site", "get_env_off_left_site", "put_env_on_right_site", "get_env_off_right_site"],
              dict(np=np, null_space=null_space, cirq=cirq))
    A = rng.normal(size=(5, 2, 2, 2)) + 1j * rng.normal(size=(5, 2, 2, 2))
    B = rng.normal(size=(5, 2, 2, 2)) + 1j * rng.normal(size=(5, 2, 2, 2))
    out["merge_A"], out["merge_B"] = A, B
    out["merge_out"] = np.stack([tet["merge"](a, b) for a, b in zip(A, B)])
    q = rng.normal(size=(5, 2, 2)) + 1j * rng.normal(size=(5, 2, 2))
    out["env_q"] = q
    L = [tet["put_env_on_left_site"](x, ret_n=True) for x in q]
    Rr = [tet["put_env_on_right_site"](x, ret_n=True) for x in q]
    out["left_U"] = np.stack([u for u, _ in L]); out["left_n"] = np.array([n for _, n in L])
    out["right_U"] = np.stack([u for u, _ in Rr]); out["right_n"] = np.array([n for _, n in Rr])
    out["left_off"] = np.stack([tet["get_env_off_left_site"](u) for u, _ in L])
    out["right_off"] = np.stack([tet["get_env_off_right_site"](u) for u, _ in Rr])
    # ---- Hamiltonian
    S = {'I': np.eye(2), 'X': X, 'Y': Y, 'Z': Z}
    gs = cut("qmps/ground_state.py", ["Hamiltonian"],
             dict(np=np, zeros=np.zeros, kron=np.kron, trace=np.trace, reduce=reduce, product=product, S=S))
    Ham = gs["Hamiltonian"]
    out["H_tfim"] = Ham({'ZZ': -1, 'X': 0.7}).to_matrix()
    out["H_heis"] = Ham({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
    out["H_mixed"] = Ham({'ZZ': -1.0, 'X': 0.3, 'IY': 0.25, 'ZI': -0.5, 'XY': 0.125}).to_matrix()
    # (Hamiltonian.from_matrix, :90-95, raises in the reference -- it takes kron of the key STRINGS -- so there
    # is nothing to record for it)
    # ---- rotosolve drivers on a deterministic 2-qubit state function (Ry Ry / CNOT / Rz Rz layers)
    CN = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)

    def state_function(p, *args):
        psi = np.array([1, 0, 0, 0], dtype=complex)
        for k in range(0, len(p), 4):
            U = np.kron(expm(-0.5j * p[k] * Y), expm(-0.5j * p[k + 1] * Y))
            V = np.kron(expm(-0.5j * p[k + 2] * Z), expm(-0.5j * p[k + 3] * X))
            psi = V @ CN @ U @ psi
        return psi
    rs = cut("qmps/rotosolve.py", ["rotosolve", "double_rotosolve"],
             dict(np=np, π=np.pi, plt=_Anything(), minimize_scalar=minimize_scalar, sinusoids=lambda *a, **k: None,
                  tqdm=lambda x: x, swap=lambda: SWAP))
    H = out["H_tfim"]
    p0 = rng.normal(size=8)
    es, Shist = rs["rotosolve"](H, state_function, p0.copy(), N_iters=4)
    out["roto_p0"] = p0
    out["roto_es"] = np.array(es)
    out["roto_S"] = np.array(Shist)
    es2, p2 = rs["double_rotosolve"](H, state_function, p0.copy(), N_iters=3)
    out["droto_es"] = np.array(es2)
    out["droto_params"] = np.array(p2)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ref_misc.npz"), **out)
    print("wrote ref_misc.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
