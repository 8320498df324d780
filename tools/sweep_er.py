#!/usr/bin/env python
"""cfg 4 (D = 8 rotosolve 3-shift energies, 64 x 64 real-form direct solve): env_real_kernel variants
-- 168 registers / 6 CTAs per SM with the streaming row builder (er_wide = 0) against 255 registers /
4 CTAs per SM with the register-cached row builder (er_wide = 1).  One JSON line per point."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qmps_b200 import _lib as L, batched as B, represent as R
    from qmps_b200.ground_state import Hamiltonian
    lib = L.require_device()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    N = 65536
    rng = np.random.default_rng(3)
    theta = torch.from_numpy(rng.normal(size=(N, 24))).to(dev)
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
    H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
    flops = 8 * 2 * 8 ** 4 + (8.0 / 3.0) * 8 ** 6
    ref = None
    for cdt, tag in ((torch.complex128, "c128"), (torch.complex64, "c64")):
        for wide in ((0, 3, 0, 3) if tag == 'c128' else (0, 2)):
            lib.qmps_set_option(b"er_wide", wide)
            fn = lambda: B.energy_theta(prog, theta, H, coord=5, shifts=B.ROTO3_SHIFTS, dtype=cdt)
            e = fn()
            e = e[0] if isinstance(e, tuple) else e
            torch.cuda.synchronize()
            if ref is None:
                ref = e.double().clone()
            err = float((e.double() - ref).abs().max())
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            print(json.dumps({"dtype": tag, "er_wide": wide, "ms": round(ms, 3), "evals_per_s": 3 * N / ms * 1e3,
                              "algo_tflops": 3 * N * flops / ms * 1e3 / 1e12, "max_abs_diff_vs_first": err}), flush=True)
    lib.qmps_set_option(b"er_wide", -1)


if __name__ == "__main__":
    main()
