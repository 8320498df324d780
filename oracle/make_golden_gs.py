"""Generate tests/golden/ref_ground_state_script.npz from the REFERENCE's own cirq-free energy route,
``scripts/ground_state_finding.py:74-128``:  ``ansatz(p)`` (Rx Rx / Rz Rz / CNOT layers, :83-92),
``state(p)`` (U = ansatz(p), V = get_env_exact(U), psi = (U x 1 x 1)(1 x U x 1)(1 x 1 x V)|0000>, :119-122),
``Ha(lambda)`` (:124-125) and ``eps(p, lambda)`` = Re <psi| 1 x Ha x 1 |psi> (:127-128).

Run in the build container only (reads /root/reference):  ``python oracle/make_golden_gs.py``.

The script imports xmps.spin / tenpy / matplotlib at module top; they are replaced by stubs.  Two of the
stubs carry arithmetic and are therefore STATED here, not taken from the reference: ``xmps.spin.paulis(0.5)``
= the Pauli matrices (X, Y, Z) and ``CNOT()`` = the standard gate with the first qubit as control
(what the call sites require: ``Rx = expm(-i theta X / 2)`` at :74-81).  ``get_env_exact`` is the reference's
own (qmps/tools.py:176-182) with xmps' ``TransferMatrix`` supplied by the oracle, as in make_golden.py.
Everything else -- the layer order, the state construction, the Hamiltonian and the energy -- is the
reference's unmodified code.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


class _Anything:
    """Permissive stand-in for plotting / tenpy objects touched at import time."""
    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_script():
    import make_golden as MG
    ref_tools = MG.load_reference_tools()                     # the reference's qmps/tools.py under stubs
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
    Z = np.array([[1, 0], [0, -1]], dtype=complex)
    CN = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    _stub("xmps")
    _stub("xmps.spin", paulis=lambda s: (X, Y, Z), CNOT=lambda: CN, swap=None, H=None, CZ=None, CRy=None)
    q = _stub("qmps")
    q.tools = ref_tools
    sys.modules["qmps.tools"] = ref_tools
    plt = _stub("matplotlib.pyplot", style=_Anything())
    _stub("matplotlib", pyplot=plt)
    for name in ("tenpy", "tenpy.networks", "tenpy.networks.mps", "tenpy.models", "tenpy.models.tf_ising",
                 "tenpy.models.spins", "tenpy.algorithms"):
        _stub(name, MPS=None, TFIChain=None, SpinModel=None, dmrg=None)
    src = open(os.path.join(REF, "scripts", "ground_state_finding.py")).read()
    # the module body below the function definitions runs optimisations and plots: keep the definitions only
    cut = src.index("def plot_convergence")
    mod = types.ModuleType("ref_gs_script")
    exec(compile(src[:cut], "ground_state_finding.py[:plot_convergence]", "exec"), mod.__dict__)
    return mod


def main():
    ref = load_script()
    rng = np.random.default_rng(11)
    out = {}
    for layers in (1, 2, 4):
        P = 4 * layers
        ps = rng.normal(size=(6, P))
        out[f"p_L{layers}"] = ps
        out[f"U_L{layers}"] = np.stack([ref.ansatz(p) for p in ps])
        for lam in (0.5, 1.0):
            out[f"eps_L{layers}_lam{lam}"] = np.array([ref.ϵ(p, lam) for p in ps])
    out["Ha_1.0"] = ref.Ha(1.0)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ref_ground_state_script.npz"), **out)
    print("wrote ref_ground_state_script.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
