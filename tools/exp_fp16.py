#!/usr/bin/env python
"""cfg 3 tile: generic fixed_point_kernel vs the register-resident D=4 kernel(s); sweeps per problem."""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from scipy.linalg import expm
    from qmps_b200 import _lib as L, batched as B, represent as R
    from qmps_b200.ground_state import Hamiltonian
    lib = L.require_device()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    NP, NT = int(os.environ.get("NP", 1024)), int(os.environ.get("NT", 128))
    rng = np.random.default_rng(2)
    theta = torch.from_numpy(rng.normal(size=(NP, 12))).to(dev)
    prog = R.ShallowCNOTStateTensor_nonuniform(4, np.zeros(12)).program()
    H = Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix()
    Wn = np.stack([expm(-1j * H * 0.02 * k * (1000 // NT)) for k in range(NT)])
    cnt = (ctypes.c_ulonglong * 4)()
    ref = None
    modes = [int(x) for x in os.environ.get("MODES", "0,6,7").split(",")]
    for cdt, tag in ((torch.complex128, "c128"), (torch.complex64, "c64")):
        A0 = B.ansatz_tensors(prog, theta[:1], dtype=cdt)[0]
        W = torch.from_numpy(Wn).to(dev).to(cdt)
        for fast in modes:
            lib.qmps_set_option(b"fp16_fast", fast)
            fn = lambda: B.loschmidt_costs(prog, theta, A0, W, dtype=cdt)
            lib.qmps_debug_counters(cnt, 1)
            c = fn()[0]
            torch.cuda.synchronize()
            lib.qmps_debug_counters(cnt, 1)
            if ref is None:
                ref = c.double().clone()
            err = float((c.double() - ref).abs().max())
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            print(json.dumps({"dtype": tag, "fp16_fast": fast, "ms": round(ms, 3), "steps_per_s": NP * NT / ms * 1e3,
                              "max_abs_diff_vs_generic_c128": err,
                              "sweeps_per_problem": cnt[1] / max(cnt[0], 1), "forced": int(cnt[2])}), flush=True)
    lib.qmps_set_option(b"fp16_fast", 0)


if __name__ == "__main__":
    main()
