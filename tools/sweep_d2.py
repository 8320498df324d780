#!/usr/bin/env python
"""D = 2 streaming kernel: time per launch vs batch size and launch options, to separate the fixed
per-launch cost from the streaming rate (t = t0 + N * 208 B / BW).  One JSON line per point."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from qmps_b200 import _lib as L
    import bench
    lib = L.require_device()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    for logn in (20, 22, 24):
        N = 1 << logn
        nbuf = 4 if logn <= 22 else 2
        A = [bench.make_tensors(torch, N, 7 + b, dev) for b in range(nbuf)]
        eta = torch.empty((N,), dtype=torch.complex128, device=dev)
        r = torch.empty((N, 2, 2), dtype=torch.complex128, device=dev)
        for pdl in (1, 0):
            for cps in (0, 1):
                lib.qmps_set_option(b"d2_pdl", pdl)
                lib.qmps_set_option(b"d2_ctas_per_sm", cps)
                steps = 200 if logn == 20 else 50 if logn == 22 else 20
                def run(k):
                    for i in range(k):
                        lib.qmps_env_exact(2, 2, N, A[i % nbuf].data_ptr(), 0, 1, eta.data_ptr(), r.data_ptr(), None, None, L.C128, st)
                run(5)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(steps); e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / steps * 1e3
                print(json.dumps({"logN": logn, "pdl": pdl, "ctas_per_sm": cps or "occ", "us_per_launch": round(us, 2),
                                  "GBps": round(208.0 * N / us / 1e3, 1)}))
        del A, eta, r
    lib.qmps_set_option(b"d2_pdl", 1); lib.qmps_set_option(b"d2_ctas_per_sm", 0)


if __name__ == "__main__":
    main()
