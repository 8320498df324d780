"""The bench kernel with all four outputs (276 B/solve): CTAs per SM x PDL sweep, rotating buffer sets, graph-free direct launches."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qmps_b200 import _lib as L
lib = L.require_device(); dev = torch.device("cuda", 0)
N, NB = 1 << 20, 4
g = torch.Generator(device=dev).manual_seed(1)
A = []
for b in range(NB):
    Z = torch.randn((N, 4, 2), dtype=torch.float64, device=dev, generator=g) + 1j * torch.randn((N, 4, 2), dtype=torch.float64, device=dev, generator=g)
    Q, _ = torch.linalg.qr(Z)
    A.append(Q.reshape(N, 2, 2, 2).permute(0, 2, 1, 3).contiguous())
outs = [(torch.empty((N,), dtype=torch.complex128, device=dev), torch.empty((N, 2, 2), dtype=torch.complex128, device=dev),
         torch.empty((N, 2, 2), dtype=torch.complex128, device=dev), torch.empty((N,), dtype=torch.int32, device=dev)) for _ in range(NB)]
st = torch.cuda.current_stream().cuda_stream
argv = [(2, 2, N, A[b].data_ptr(), 0, 1, outs[b][0].data_ptr(), outs[b][1].data_ptr(), outs[b][2].data_ptr(), outs[b][3].data_ptr(), L.C128, st) for b in range(NB)]
for ctas in (0, 1, 2, 3):
    for pdl in (1, 0):
        lib.qmps_set_option(b"d2_ctas_per_sm", ctas); lib.qmps_set_option(b"d2_pdl", pdl)
        for i in range(50):
            lib.qmps_env_exact(*argv[i % NB])
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                lib.qmps_env_exact(*argv[i % NB])
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 200)
        print(json.dumps({"ctas_per_sm": ctas, "pdl": pdl, "us_per_launch": best * 1e3, "gbs": 276 * N / (best * 1e-3) / 1e9}), flush=True)
lib.qmps_set_option(b"d2_ctas_per_sm", 1); lib.qmps_set_option(b"d2_pdl", 1)
