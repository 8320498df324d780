"""Bring-up check of the tcgen05 complex64 path (qmps_cgemm_c64_tc, qmps_tm_power c64) against numpy.
Run on the GPU box:  timeout 300 python tools/tc_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qmps_b200 import _lib as L, batched  # noqa: E402
import oracle as O  # noqa: E402  (checker only)


def cgemm(lib, X, Y, conj):
    batch, nsum, M, K = X.shape
    N = Y.shape[2]
    C = torch.empty((batch, M, N), dtype=torch.complex64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.qmps_cgemm_c64_tc(batch, nsum, M, N, K, X.data_ptr(), Y.data_ptr(), int(conj), C.data_ptr(), st), "cgemm")
    torch.cuda.synchronize()
    return C


def case(lib, batch, nsum, M, N, K, conj, seed, verbose=False):
    rng = np.random.default_rng(seed)
    X = (rng.normal(size=(batch, nsum, M, K)) + 1j * rng.normal(size=(batch, nsum, M, K))).astype(np.complex64)
    Y = (rng.normal(size=(batch, nsum, N, K)) + 1j * rng.normal(size=(batch, nsum, N, K))).astype(np.complex64)
    ref = np.einsum("btmk,btnk->bmn", X.astype(np.complex128), (Y.conj() if conj else Y).astype(np.complex128))
    got = cgemm(lib, torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda(), conj).cpu().numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"cgemm batch={batch} nsum={nsum} M={M} N={N} K={K} conj={conj}: rel err {err:.3e}", flush=True)
    if verbose or err > 1e-5:
        e = np.abs(got[0] - ref[0])
        blk = e.reshape(M // 8, 8, N // 8, 8).max(axis=(1, 3))
        np.set_printoptions(linewidth=250, precision=1)
        print("per 8x8 block max error of matrix 0 (rows = C row blocks):")
        print(blk[:16, :16])
        print("got[0,0,:4]", got[0, 0, :4], "ref", ref[0, 0, :4])
        print("got[0,1,:4]", got[0, 1, :4], "ref", ref[0, 1, :4])
        # hypotheses: real / imaginary parts separately
        print("re err", np.abs(got[0].real - ref[0].real).max(), "im err", np.abs(got[0].imag - ref[0].imag).max())
    return err


def main():
    lib = L.require_device()
    ok = True
    for persistent in (0, 1):
        lib.qmps_set_option(b"tc_persistent", persistent)
        print("== tc_persistent", persistent, flush=True)
        ok &= case(lib, 1, 1, 64, 64, 32, 0, 1) < 1e-5
        ok &= case(lib, 1, 1, 64, 64, 32, 1, 2) < 1e-5
        ok &= case(lib, 1, 1, 64, 64, 64, 0, 3) < 1e-5
        ok &= case(lib, 2, 2, 64, 64, 128, 1, 4) < 1e-5
        ok &= case(lib, 3, 1, 128, 192, 96, 0, 5) < 1e-5
        ok &= case(lib, 40, 2, 128, 128, 64, 1, 6) < 1e-5      # 160 tiles > 148 SMs: two tiles on some CTAs
        ok &= case(lib, 700, 1, 64, 64, 64, 0, 7) < 1e-5       # many tiles per CTA, ring wraps
    lib.qmps_set_option(b"tc_persistent", 1)
    # power method
    from scipy.stats import unitary_group
    for D, cnt in ((64, 3), (128, 2), (256, 2)):
        def tensors(seed):
            out = []
            for k in range(cnt):
                U = unitary_group.rvs(2 * D, random_state=seed + k)
                out.append(O.unitary_to_tensor(U))
            return np.stack(out)
        A, B = tensors(400 + D), tensors(500 + D)
        for tcp in (1, 0):
            lib.qmps_set_option(b"tc_power", tcp)
            r, ray = batched.tm_power(torch.from_numpy(A).cuda().to(torch.complex64), torch.from_numpy(B).cuda().to(torch.complex64), K=8)
            torch.cuda.synchronize()
            e1 = e2 = 0.0
            for k in range(cnt):
                r0, q0 = O.power_method(A[k], B[k], 8)
                e1 = max(e1, np.abs(r[k].cpu().numpy() - r0).max())
                e2 = max(e2, abs(ray[k].item() - q0))
            print(f"tm_power c64 D={D} tc_power={tcp}: max |r - r_oracle| {e1:.3e}  |rayleigh - oracle| {e2:.3e}", flush=True)
            if tcp:
                ok &= e1 < 1e-5 and e2 < 1e-5
    lib.qmps_set_option(b"tc_power", 1)
    # timing
    for D, N in ((64, 512), (256, 32)):
        g = torch.Generator(device="cuda").manual_seed(D)
        A = torch.randn((N, 2, D, D, 2), device="cuda", generator=g)
        A = torch.view_as_complex(A).to(torch.complex64) / np.sqrt(2 * D)
        Bm = torch.view_as_complex(torch.randn((N, 2, D, D, 2), device="cuda", generator=g)).to(torch.complex64) / np.sqrt(2 * D)
        for tcp in (1, 0):
            lib.qmps_set_option(b"tc_power", tcp)
            for rep in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                batched.tm_power(A, Bm, K=32)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print(f"tm_power c64 D={D} N={N} K=32 tc_power={tcp}: {ms:.3f} ms  -> {33 * N * 32.0 * D ** 3 / ms / 1e9:.1f} algorithmic TFLOP/s", flush=True)
    lib.qmps_set_option(b"tc_power", 1)
    print("TC CHECK", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
