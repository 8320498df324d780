#!/usr/bin/env python
"""cfg 3 tile (16 x 16 mixed two-site maps, all eigenvalues): lanes per problem x CTA size sweep of
fixed_point_kernel, complex128 and complex64.  One JSON line per point."""
import json, os, sys
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
    NP, NT = 1024, 128
    rng = np.random.default_rng(2)
    theta = torch.from_numpy(rng.normal(size=(NP, 12))).to(dev)
    prog = R.ShallowCNOTStateTensor_nonuniform(4, np.zeros(12)).program()
    H = Hamiltonian({'ZZ': -1, 'X': 0.2}).to_matrix()
    Wn = np.stack([expm(-1j * H * 0.04 * k) for k in range(NT)])
    ref = None
    for cdt, tag in ((torch.complex128, "c128"), (torch.complex64, "c64")):
        A0 = B.ansatz_tensors(prog, theta[:1], dtype=cdt)[0]
        W = torch.from_numpy(Wn).to(dev).to(cdt)
        for grp, blk in ((16, 128), (16, 64), (8, 128), (8, 64), (4, 128), (4, 64), (4, 32)):
            lib.qmps_set_option(b"fp_group", grp)
            lib.qmps_set_option(b"fp_block", blk)
            fn = lambda: B.loschmidt_costs(prog, theta, A0, W, dtype=cdt)
            c = fn()[0]
            torch.cuda.synchronize()
            if tag == "c128":
                if ref is None:
                    ref = c.clone()
                err = float((c - ref).abs().max())
            else:
                err = float((c.double() - ref).abs().max())
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            print(json.dumps({"dtype": tag, "lanes": grp, "block": blk, "ms": round(ms, 3),
                              "steps_per_s": NP * NT / ms * 1e3, "max_abs_diff_vs_default": err}), flush=True)
    lib.qmps_set_option(b"fp_group", 0); lib.qmps_set_option(b"fp_block", 0)


if __name__ == "__main__":
    main()
