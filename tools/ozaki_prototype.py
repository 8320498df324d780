#!/usr/bin/env python
"""Accuracy prototype (numpy, CPU) for the planned INT8-slice (Ozaki-style) complex128 contraction on
`tcgen05 kind::i8` -- NOT a product path, only the arithmetic the kernel would do:

  * every real matrix is scaled per row (left operand) / per column (right operand) by a power of two so that
    |x| < 1, then cut into `s` signed slices of `b` bits each (x = sum_k q_k 2^{-b k}, |q_k| < 2^{b-1}: int8 for b <= 7);
  * slice products q_i^A . q_j^B are EXACT in int32 for K <= 2^31 / 2^(2b-2); only pairs with i + j < s are formed
    (the rest is below the target precision) and pairs with equal i + j are added in integers before one conversion;
  * the complex product is four real products.

Prints, for one application r' = sum_s A_s r B_s^dagger at D = 64 / 256, the relative error against numpy's complex128
result and the number of int8 GEMMs per real GEMM, for several (bits, slices).

  python tools/ozaki_prototype.py
"""
import json

import numpy as np


def slices(x, axis, b, s):
    """x (real matrix) -> (scale exponents along `axis` kept, int slices [s, ...]) with x ~ 2^e * sum_k q_k 2^{-b(k+1)}."""
    amax = np.max(np.abs(x), axis=axis, keepdims=True)
    e = np.ceil(np.log2(np.maximum(amax, 1e-300))) + 1                 # |x| / 2^e < 1/2
    y = x / np.exp2(e)
    qs = []
    for _ in range(s):
        y = y * (1 << b)
        q = np.rint(y)                                                 # |q| <= 2^(b-1)
        qs.append(q.astype(np.int64))
        y = y - q
    return e, np.stack(qs)


def real_gemm(A, B, b, s):
    """A (M x K) . B (K x N) from int slices; returns (result, number of integer GEMMs)."""
    ea, qa = slices(A, 1, b, s)
    eb, qb = slices(B, 0, b, s)
    acc = np.zeros((A.shape[0], B.shape[1]))
    n = 0
    for t in range(s):                                                 # t = i + j
        g = np.zeros((A.shape[0], B.shape[1]), dtype=np.int64)
        for i in range(t + 1):
            g += qa[i] @ qb[t - i]                                     # exact (int32 range holds for b <= 7, K <= 2^17)
            n += 1
        acc += g.astype(np.float64) * 2.0 ** (-b * (t + 2))
    return acc * np.exp2(ea) * np.exp2(eb), n


def cgemm(X, Y, b, s):
    rr, n = real_gemm(X.real, Y.real, b, s)
    ii, _ = real_gemm(X.imag, Y.imag, b, s)
    ri, _ = real_gemm(X.real, Y.imag, b, s)
    ir, _ = real_gemm(X.imag, Y.real, b, s)
    return (rr - ii) + 1j * (ri + ir), n


def main():
    rng = np.random.default_rng(0)
    for D in (64, 256):
        Z = rng.normal(size=(2 * D, D)) + 1j * rng.normal(size=(2 * D, D))
        A = np.linalg.qr(Z)[0].reshape(D, 2, D).transpose(1, 0, 2)     # left-canonical A[s, i, j]
        Z = rng.normal(size=(2 * D, D)) + 1j * rng.normal(size=(2 * D, D))
        B = np.linalg.qr(Z)[0].reshape(D, 2, D).transpose(1, 0, 2)
        r = rng.normal(size=(D, D)) + 1j * rng.normal(size=(D, D))
        r /= np.linalg.norm(r)
        ref = sum(A[k] @ r @ B[k].conj().T for k in range(2))
        for b, s in ((7, 4), (7, 5), (7, 6), (7, 7), (6, 7), (6, 8)):
            out = np.zeros_like(ref)
            n = 0
            for k in range(2):
                T, n1 = cgemm(A[k], r, b, s)
                P, n2 = cgemm(T, B[k].conj().T, b, s)
                out += P
                n = n1
            err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
            print(json.dumps({"D": D, "bits": b, "slices": s, "int8_gemms_per_real_gemm": n, "rel_err": float(err)}))


if __name__ == "__main__":
    main()
