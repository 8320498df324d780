"""Generate tests/golden/ref_*.npz from the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):  ``python oracle/make_golden.py``.

``qmps/tools.py`` imports xmps / cirq / matplotlib / tqdm at module top; those
packages are absent here, so they are replaced by empty stub modules.  The
functions recorded below do not touch the stubs -- they are the reference's
unmodified numpy/scipy code.  ``get_env_exact`` additionally needs
``TransferMatrix`` (xmps, un-vendored): it is supplied by the oracle's
restatement, so that golden pins everything in the chain EXCEPT the eigen-solve
(unitary_to_tensor -> [eigs] -> cholesky -> environment_to_unitary).
"""
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
from scipy.stats import unitary_group

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_tools():
    import oracle

    class TransferMatrix:                       # xmps stand-in (oracle restatement)
        def __init__(self, A):
            self.A = A

        def eigs(self):
            return oracle.eigs(self.A)

    _stub("xmps")
    _stub("xmps.spin", U4=None)
    _stub("xmps.iMPS", TransferMatrix=TransferMatrix, iMPS=None)
    _stub("cirq")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except ImportError:
            _stub("tqdm", tqdm=lambda x, *a, **k: x)
    spec = importlib.util.spec_from_file_location("ref_qmps_tools", os.path.join(REF, "qmps", "tools.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_exact_loschmidt():
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    spec = importlib.util.spec_from_file_location(
        "ref_exact_loschmidt", os.path.join(REF, "qmps", "loschmidts", "exact_loschmidt.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def haar(n, seed):
    return unitary_group.rvs(n, random_state=seed)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference_tools()
    out = {}

    # --- unitary_to_tensor / tensor_to_unitary on Haar unitaries, D = 2, 4, 8
    for D in (2, 4, 8):
        Us = np.stack([haar(2 * D, 100 * D + k) for k in range(4)])
        As = np.stack([ref.unitary_to_tensor(U) for U in Us])
        U2 = np.stack([ref.tensor_to_unitary(A) for A in As])
        out[f"u2t_U_D{D}"] = Us
        out[f"u2t_A_D{D}"] = As
        out[f"t2u_U_D{D}"] = U2
    U, passed = ref.tensor_to_unitary(out["u2t_A_D2"][0], testing=True)
    out["t2u_testing_passed"] = np.array(bool(passed))

    # --- unitary_extension: tall, wide, padded
    rng = np.random.default_rng(7)
    Qt = np.linalg.qr(rng.normal(size=(6, 3)) + 1j * rng.normal(size=(6, 3)))[0]
    out["uext_tall_in"] = Qt
    out["uext_tall_out"] = ref.unitary_extension(Qt)
    out["uext_wide_in"] = Qt.conj().T
    out["uext_wide_out"] = ref.unitary_extension(Qt.conj().T)
    out["uext_pad_out"] = ref.unitary_extension(Qt, 8)

    # --- environment_to_unitary / from_unitary
    C = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    V = ref.environment_to_unitary(C)
    out["e2u_in"] = C
    out["e2u_out"] = V
    out["efu_out"] = ref.environment_from_unitary(V)
    C4 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    out["e2u_in_D4"] = C4
    out["e2u_out_D4"] = ref.environment_to_unitary(C4)

    # --- real/complex packing, cT, direct_sum
    v = rng.normal(size=10)
    out["frv_in"] = v
    out["frv_out"] = ref.from_real_vector(v)
    out["trv_out"] = ref.to_real_vector(C)
    T = rng.normal(size=(3, 2, 4)) + 1j * rng.normal(size=(3, 2, 4))
    out["cT_in"] = T
    out["cT_out"] = ref.cT(T)
    out["dsum_out"] = ref.direct_sum(np.real(C), np.eye(3))

    # --- get_env_exact (reference chain, oracle eigen-solve), D = 2, 4
    for D in (2, 4):
        Us = out[f"u2t_U_D{D}"]
        out[f"env_V_D{D}"] = np.stack([ref.get_env_exact(U) for U in Us])

    # --- double_rotosolve on a deterministic two-frequency cost
    def eps(p):
        return (np.sin(p[0]) * np.cos(2 * p[1]) + 0.3 * np.sin(2 * p[0] + 0.4)
                + 0.5 * np.cos(p[1] - 0.2) + 0.1 * np.sin(p[2]) * np.sin(p[0]))
    p0 = np.array([0.3, -1.1, 2.0])
    res = ref.double_rotosolve(eps, p0.copy(), 3, False)
    out["drs_p0"] = p0
    out["drs_history"] = np.array(res.history)
    out["drs_x"] = np.array(res.x)

    np.savez(os.path.join(OUT, "ref_tools.npz"), **out)
    print("wrote", os.path.join(OUT, "ref_tools.npz"), len(out), "arrays")

    # --- exact Loschmidt rate function (reference scipy code)
    el = load_reference_exact_loschmidt()
    ts = np.array([0.0, 0.5, 1.0, 2.0, 3.0, 4.5, 6.0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vals = np.array([el.loschmidt(t, 1.5, 0.2) for t in ts])
        vals2 = np.array([el.loschmidt(t, 0.5, 2.0) for t in ts])
    np.savez(os.path.join(OUT, "ref_exact_loschmidt.npz"), t=ts, g15_02=vals, g05_20=vals2)
    print("wrote ref_exact_loschmidt.npz", vals)

    # --- fixtures/A.npy (the reference's one stored tensor) as a fixed INPUT
    raw = np.load(os.path.join(REF, "fixtures", "A.npy"))
    d, D = int(raw[0].real), int(raw[1].real)
    np.savez(os.path.join(OUT, "ref_fixture_A.npz"), raw=raw, A=raw[3:].reshape(d, D, D))
    print("wrote ref_fixture_A.npz")


if __name__ == "__main__":
    main()
