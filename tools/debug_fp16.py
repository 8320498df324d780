#!/usr/bin/env python
"""Correctness + sweep statistics of the D = 4 QR kernel on a small batch (debug aid)."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qmps_b200 import batched as B, _lib as L
lib = L.require_device()
rng = np.random.default_rng(0)
N = 4096
def lc(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    Z = torch.view_as_complex(torch.randn((n, 8, 4, 2), dtype=torch.float64, device="cuda", generator=g))
    Q, _ = torch.linalg.qr(Z)
    return Q.reshape(n, 4, 2, 4).permute(0, 2, 1, 3).contiguous()
A, Bt = lc(N, 1), lc(N, 2)
cnt = (ctypes.c_ulonglong * 4)()
for left in (False, True):
    lib.qmps_debug_counters(cnt, 1)
    lib.qmps_set_option(b"fp16_fast", 1)
    t0 = time.time(); fast = B.fixed_point(A, Bt, left=left, want_vec=False); torch.cuda.synchronize(); t1 = time.time()
    lib.qmps_debug_counters(cnt, 1)
    lib.qmps_set_option(b"fp16_fast", 0)
    slow = B.fixed_point(A, Bt, left=left, want_vec=False); torch.cuda.synchronize()
    E = torch.einsum("nsij,nskl->nikjl", A, Bt.conj()).reshape(N, 16, 16)
    w = torch.linalg.eigvals(E).abs().max(dim=1).values
    print("left", left, "problems", cnt[0], "sweeps/problem", cnt[1] / max(cnt[0], 1), "forced", cnt[2],
          "max|fast-slow|", (fast.eta.abs() - slow.eta.abs()).abs().max().item(),
          "max|fast-eig|", (fast.eta.abs() - w).abs().max().item(), "status sum", int(fast.status.sum()), "ms", (t1 - t0) * 1e3)
