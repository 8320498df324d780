"""Generate tests/golden/ref_brickwall.npz from the REFERENCE's own brick-wall code.

Run in the build container only (reads /root/reference):  ``python oracle/make_golden_bw.py``.

``new_tdvp/ClassicalTDVPStripped.py`` imports jax / xmps / cirq / matplotlib / qmps at module
top; those are replaced by empty stub modules.  Everything recorded below is the reference's
unmodified numpy/scipy code (``np.einsum`` contractions + ``scipy.linalg.eig``) called on
seeded Haar unitaries: ``RightEnvironment`` / ``LeftEnvironment`` (``exact_environment_circuit``,
``exact_environment``, ``circuit``), ``OverlapCalculator`` (2- and 4-qubit expectation values, both
the einsum and the matrix forms), ``ManifoldOverlap`` (``circuit`` and ``mcircuit``) and the body of
``Evolve.exact_cost_function`` (``:777-790``) with the candidate unitaries given directly --
``paramU`` (``:159-180``) needs ``xmps.spin.U4``, which is not vendored, so the parameter -> unitary
step is NOT pinned and the batched API takes unitaries.
"""
import importlib.util
import os
import sys
import types

import numpy as np
from scipy.stats import unitary_group

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_bw():
    _stub("jax", device_put=None, jit=lambda f, *a, **k: f)
    _stub("jax.numpy")
    _stub("xmps")
    _stub("xmps.spin", U4=None, lambdas=None)
    _stub("qmps")
    _stub("qmps.ground_state", Hamiltonian=None)
    _stub("qmps.represent", ShallowFullStateTensor=None)
    _stub("cirq")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    spec = importlib.util.spec_from_file_location(
        "ref_bw", os.path.join(REF, "new_tdvp", "ClassicalTDVPStripped.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_bw()
    RE, LE, OC, MO = ref.RightEnvironment(), ref.LeftEnvironment(), ref.OverlapCalculator(), ref.ManifoldOverlap()
    N = 24
    rs = np.random.RandomState(20201)
    U1 = np.stack([unitary_group.rvs(4, random_state=rs) for _ in range(N)])
    U2 = np.stack([unitary_group.rvs(4, random_state=rs) for _ in range(N)])
    V1 = np.stack([unitary_group.rvs(4, random_state=rs) for _ in range(N)])     # candidate state U1', U2'
    V2 = np.stack([unitary_group.rvs(4, random_state=rs) for _ in range(N)])
    # a few "close" candidates (what the TDVP optimiser actually visits): V = U . exp(small)
    from scipy.linalg import expm
    for k in range(N // 2):
        h1 = rs.randn(4, 4) + 1j * rs.randn(4, 4)
        h2 = rs.randn(4, 4) + 1j * rs.randn(4, 4)
        V1[k] = U1[k] @ expm(0.05j * (h1 + h1.conj().T))
        V2[k] = U2[k] @ expm(0.05j * (h2 + h2.conj().T))
    M = np.stack([unitary_group.rvs(2, random_state=rs) for _ in range(N)])
    O2 = rs.randn(N, 4, 4) + 1j * rs.randn(N, 4, 4)
    O2 = O2 + O2.conj().transpose(0, 2, 1)
    O4 = rs.randn(N, 16, 16) + 1j * rs.randn(N, 16, 16)
    O4 = O4 + O4.conj().transpose(0, 2, 1)
    Hh = rs.randn(16, 16) + 1j * rs.randn(16, 16)
    W = expm(-0.1j * (Hh + Hh.conj().T))

    out = dict(U1=U1, U2=U2, V1=V1, V2=V2, M=M, O2=O2, O4=O4, W=W)
    keys = ("renv_mat", "lenv_mat", "renv_eta", "renv_vec", "lenv_eta", "lenv_vec", "renv_circuit",
            "renv_mat_same", "renv_eta_same", "renv_vec_same", "exp2", "mexp2", "exp4", "mexp4",
            "overlap", "moverlap", "exact_cost")
    acc = {k: [] for k in keys}
    for k in range(N):
        u1, u2 = U1[k].reshape(2, 2, 2, 2), U2[k].reshape(2, 2, 2, 2)
        v1_, v2_ = V1[k].conj().T.reshape(2, 2, 2, 2), V2[k].conj().T.reshape(2, 2, 2, 2)
        u1_, u2_ = U1[k].conj().T.reshape(2, 2, 2, 2), U2[k].conj().T.reshape(2, 2, 2, 2)
        acc["renv_mat"].append(RE.exact_environment_circuit(u1, u2, v1_, v2_))
        acc["lenv_mat"].append(LE.exact_environment_circuit(u1, u2, v1_, v2_))
        e, v = RE.exact_environment(u1, u2, v1_, v2_)
        acc["renv_eta"].append(e); acc["renv_vec"].append(v)
        e, v = LE.exact_environment(u1, u2, v1_, v2_)
        acc["lenv_eta"].append(e); acc["lenv_vec"].append(v)
        acc["renv_circuit"].append(RE.circuit(u1, u2, v1_, v2_, M[k]))
        acc["renv_mat_same"].append(RE.exact_environment_circuit(u1, u2, u1_, u2_))
        e, v = RE.exact_environment(u1, u2, u1_, u2_)
        acc["renv_eta_same"].append(e); acc["renv_vec_same"].append(v)
        acc["exp2"].append(OC.expectation_value(u1, u2, O2[k].reshape(2, 2, 2, 2)))
        acc["mexp2"].append(OC.mexpectation_value(U1[k], U2[k], O2[k]))
        acc["exp4"].append(OC.expectation_value(u1, u2, O4[k].reshape((2,) * 8)))
        acc["mexp4"].append(OC.mexpectation_value(U1[k], U2[k], O4[k]))
        # Evolve.exact_cost_function body (ClassicalTDVPStripped.py:777-790), U1_, U2_ given
        Mr, Ml = ref.Represent.exact_env(types.SimpleNamespace(RE=RE, LE=LE), u1, u2, v1_, v2_)
        ov = MO.circuit(u1, u2, v1_, v2_, Mr, Mr.conj().T, W.reshape((2,) * 8), "greedy")
        acc["overlap"].append(ov)
        acc["exact_cost"].append(-np.abs(ov) ** 2)
        acc["moverlap"].append(MO.mcircuit(U1[k], U2[k], V1[k].conj().T, V2[k].conj().T, Mr, Mr.conj().T, W))
    for k in keys:
        out[k] = np.asarray(acc[k])
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "ref_brickwall.npz"), **out)
    print("wrote ref_brickwall.npz:", {k: v.shape for k, v in out.items()})
    print("einsum vs matrix forms: exp2 %.2e exp4 %.2e overlap %.2e" % (
        np.abs(out["exp2"] - out["mexp2"].real).max(), np.abs(out["exp4"] - out["mexp4"]).max(),
        np.abs(out["overlap"] - out["moverlap"]).max()))


if __name__ == "__main__":
    main()
