// qmps_b200 two-layer brick-wall iMPS primitives: environments, expectation
// values and the manifold overlap of new_tdvp/ClassicalTDVPStripped.py, executed by a GROUP of
// cooperating lanes on per-problem scratch (shared memory on the device, plain arrays in the
// host-emu test build).
//
// The reference writes each quantity as an np.einsum over (2,2,2,2) views of 4x4 unitaries
// (ClassicalTDVPStripped.py:228-533).  Here they are state-vector / 4x4 forms:
//   u2 = U2[:,0] (the column reached from |00>), w2 = U2_[0,:] (the bra row), P = U1_ . U1
//   right map   M[(a,b),(c,e)] = sum_xy P[(b,y),(a,x)] u2[x,c] w2[y,e]            (:394-417)
//   left  map   M[(a,b),(c,e)] = sum_xy u2[c,x] w2[e,y] P[(y,b),(x,a)]            (:322-345)
//   |psi>  = (1 (x) U1 (x) .. (x) 1)(U2 (x) .. (x) U2)|0..0>,  <phi| likewise from U1_, U2_
//   <O>     = Re <psi| 1 (x) O (x) 1 |psi>                                         (:428-533)
//   overlap = <phi| Ml (x) W (x) Mr |psi>                                          (:228-268)
// Qubit 0 is the most significant bit of a state index (numpy kron order).
#pragma once
#include "core.cuh"

namespace qmps {

template <typename T> struct BwWork {
  cx<T>* k1;    // [16] U1
  cx<T>* k2;    // [4]  U2[:,0]
  cx<T>* b1;    // [16] U1_  (as the reference passes it: already daggered)
  cx<T>* b2;    // [4]  U2_[0,:]
  cx<T>* P;     // [16] U1_ . U1
  cx<T>* E;     // [4 x 5] map / augmented system
  cx<T>* w;     // [4] eigenvalues
  cx<T>* vv;    // [4]
  cx<T>* rc;    // [4]
  cx<T>* rs;    // [4]
  cx<T>* x;     // [4] eigenvector
  cx<T>* ml;    // [4] Ml
  cx<T>* mr;    // [4] Mr
  T* rn;        // [4]
  int* step;    // [4]
  int* done;    // [4]
  T* red;       // [g.size]
  cx<T>* psi;   // [64]
  cx<T>* phi;   // [64]
  cx<T>* tmp;   // [64]
};
constexpr int BW_CX_ELEMS = 16 + 4 + 16 + 4 + 16 + 20 + 4 * 7 + 3 * 64;     // complex scratch per problem

template <typename T> QMPS_HD BwWork<T> bw_carve(unsigned char* base, int gsize) {
  BwWork<T> W;
  cx<T>* c = reinterpret_cast<cx<T>*>(base);
  W.k1 = c; c += 16; W.k2 = c; c += 4; W.b1 = c; c += 16; W.b2 = c; c += 4; W.P = c; c += 16;
  W.E = c; c += 20; W.w = c; c += 4; W.vv = c; c += 4; W.rc = c; c += 4; W.rs = c; c += 4;
  W.x = c; c += 4; W.ml = c; c += 4; W.mr = c; c += 4;
  W.psi = c; c += 64; W.phi = c; c += 64; W.tmp = c; c += 64;
  T* r = reinterpret_cast<T*>(c);
  W.rn = r; r += 4; W.red = r; r += gsize;
  int* ip = reinterpret_cast<int*>(r);
  W.step = ip; ip += 4; W.done = ip;
  return W;
}
template <typename T> QMPS_HD size_t bw_work_bytes(int gsize) {
  size_t b = sizeof(cx<T>) * BW_CX_ELEMS + sizeof(T) * (4 + gsize) + sizeof(int) * 8;
  return (b + 15) & ~size_t(15);
}

// inputs -> scratch.  bra_undaggered: B1, B2 are the candidate unitaries V1, V2 themselves
// (what paramU returns, :159-180) and the daggers U1_ = V1^dagger, U2_ = V2^dagger are formed here.
template <typename T>
QMPS_HDN void bw_load(const Grp& g, const cx<T>* U1, const cx<T>* U2, const cx<T>* B1, const cx<T>* B2,
                      int bra_undaggered, const BwWork<T>& W) {
  for (int e = g.lane; e < 16; e += g.size) {
    W.k1[e] = U1[e];
    if (B1) {
      const int r = e >> 2, c = e & 3;
      W.b1[e] = bra_undaggered ? conj(B1[c * 4 + r]) : B1[e];
    }
  }
  for (int q = g.lane; q < 4; q += g.size) {
    W.k2[q] = U2[q * 4];
    if (B2) W.b2[q] = bra_undaggered ? conj(B2[q * 4]) : B2[q];
  }
  g.sync();
  if (B1) {
    for (int e = g.lane; e < 16; e += g.size) {
      const int r = e >> 2, c = e & 3;
      cx<T> acc = mk<T>(0, 0);
      for (int k = 0; k < 4; ++k) cmad(acc, W.b1[r * 4 + k], W.k1[k * 4 + c]);
      W.P[e] = acc;
    }
    g.sync();
  }
}

// 4x4 environment map into E (leading dimension ld); side 0 = right, 1 = left
template <typename T>
QMPS_HDN void bw_env_matrix(const Grp& g, const BwWork<T>& W, int side, cx<T>* E, int ld) {
  for (int e = g.lane; e < 16; e += g.size) {
    const int row = e >> 2, col = e & 3;
    const int a = row >> 1, b = row & 1, c = col >> 1, ee = col & 1;
    cx<T> acc = mk<T>(0, 0);
    for (int x = 0; x < 2; ++x)
      for (int y = 0; y < 2; ++y) {
        cx<T> p, uw;
        if (side == 0) { p = W.P[(2 * b + y) * 4 + (2 * a + x)]; uw = W.k2[2 * x + c] * W.b2[2 * y + ee]; }
        else { p = W.P[(2 * y + b) * 4 + (2 * x + a)]; uw = W.k2[2 * c + x] * W.b2[2 * ee + y]; }
        cmad(acc, p, uw);
      }
    E[row * ld + col] = acc;
  }
}

// RightEnvironment.circuit (:360-384): out[i][j] = sum_xyzw w2[y,z] P[(i,y),(j,x)] M[z,w] u2[x,w]
template <typename T>
QMPS_HDN void bw_env_apply(const Grp& g, const BwWork<T>& W, const cx<T>* M, cx<T>* out) {
  for (int e = g.lane; e < 4; e += g.size) {
    const int i = e >> 1, j = e & 1;
    cx<T> acc = mk<T>(0, 0);
    for (int x = 0; x < 2; ++x)
      for (int y = 0; y < 2; ++y) {
        cx<T> s = mk<T>(0, 0);
        for (int z = 0; z < 2; ++z)
          for (int w = 0; w < 2; ++w) cmad(s, W.b2[2 * y + z] * M[2 * z + w], W.k2[2 * x + w]);
        cmad(acc, W.P[(2 * i + y) * 4 + (2 * j + x)], s);
      }
    out[e] = acc;
  }
}

// numpy's argmax on a complex array: lexicographic on (real, imag), first one on ties -- the
// reference's selection rule (:351, :423), not the largest modulus
template <typename T> QMPS_HD int bw_argmax_lex(const cx<T>* w, int n) {
  int k = 0;
  for (int i = 1; i < n; ++i)
    if (w[i].re > w[k].re || (w[i].re == w[k].re && w[i].im > w[k].im)) k = i;
  return k;
}

// exact_environment (:347-352, :419-426): all eigenvalues of the 4x4 map (Hessenberg + QR, as
// scipy.linalg.eig), numpy-argmax selection, eigenvector by inverse iteration in zgeev's gauge
// (unit 2-norm, component of largest modulus real positive).  x = W.x.  mat_out (optional): the map.
template <typename T>
QMPS_HDN int bw_exact_environment(const Grp& g, const BwWork<T>& W, int side, cx<T>* lambda_out, int want_vec) {
  const int ld = 5, n = 4;
  bw_env_matrix<T>(g, W, side, W.E, ld);
  g.sync();
  hessenberg<T>(g, W.E, ld, n, W.vv);
  const int fail = hqr_eigenvalues<T>(g, W.E, ld, n, W.w, W.rc, W.rs, W.rn);
  g.sync();
  const int kmax = bw_argmax_lex<T>(W.w, n);
  const cx<T> lam = W.w[kmax];
  *lambda_out = lam;
  const int status = fail ? ST_NO_CONVERGE : ST_OK;
  if (!want_vec) return status;
  g.sync();
  bw_env_matrix<T>(g, W, side, W.E, ld);
  g.sync();
  for (int i = g.lane; i < n; i += g.size) {
    W.E[i * ld + i] = W.E[i * ld + i] - lam;
    T t = T(0.61803398874989485) * T(i + 1);
    t -= floor(t);
    W.E[i * ld + n] = mk<T>(T(0.5) + t, T(0.25) - T(0.5) * t);
  }
  g.sync();
  T scale = cabs(lam);
  if (!(scale > T(1e-30))) scale = T(1);
  lu_solve_aug<T>(g, W.E, ld, n, W.x, W.step, W.done, eps_of<T>::v() * scale, 1);
  T nrm2 = T(0), bigv = T(-1);
  int big = 0;
  for (int i = 0; i < n; ++i) {
    const T a = norm2(W.x[i]);
    nrm2 += a;
    if (a > bigv) { bigv = a; big = i; }
  }
  const cx<T> ph = conj(W.x[big]) * (T(1) / sqrt(bigv)) * (T(1) / sqrt(nrm2));
  g.sync();
  for (int i = g.lane; i < n; i += g.size) W.x[i] = W.x[i] * ph;
  g.sync();
  return status;
}

// product state v4 (x) v4 (x) ... on 2*cells qubits
template <typename T>
QMPS_HDN void bw_product(const Grp& g, const cx<T>* v4, int cells, cx<T>* out) {
  const int n = 1 << (2 * cells);
  for (int idx = g.lane; idx < n; idx += g.size) {
    cx<T> p = v4[(idx >> (2 * (cells - 1))) & 3];
    for (int c = 1; c < cells; ++c) p = p * v4[(idx >> (2 * (cells - 1 - c))) & 3];
    out[idx] = p;
  }
}

// 4x4 matrix on the qubit pair whose 2-bit field starts at bit `sh`:
//   from_right = 0:  out = (.. U ..) in      (ket)      from_right = 1:  out = in (.. U ..)   (bra row)
template <typename T>
QMPS_HDN void bw_apply4(const Grp& g, const cx<T>* U, int n, int sh, int from_right, const cx<T>* in, cx<T>* out) {
  for (int idx = g.lane; idx < n; idx += g.size) {
    const int r = (idx >> sh) & 3, base = idx & ~(3 << sh);
    cx<T> acc = mk<T>(0, 0);
    for (int c = 0; c < 4; ++c) {
      if (from_right) cmad(acc, in[base | (c << sh)], U[c * 4 + r]);
      else cmad(acc, U[r * 4 + c], in[base | (c << sh)]);
    }
    out[idx] = acc;
  }
}

// ket / bra of `cells` unit cells: result in buf, scratch in tmp (both >= 4^cells entries)
template <typename T>
QMPS_HDN void bw_build_state(const Grp& g, const cx<T>* v4, const cx<T>* U, int cells, int from_right,
                             cx<T>* buf, cx<T>* tmp) {
  const int n = 1 << (2 * cells);
  bw_product<T>(g, v4, cells, cells == 2 ? tmp : buf);
  g.sync();
  if (cells == 2) {
    bw_apply4<T>(g, U, n, 1, from_right, tmp, buf);
  } else {
    bw_apply4<T>(g, U, n, 3, from_right, buf, tmp);
    g.sync();
    bw_apply4<T>(g, U, n, 1, from_right, tmp, buf);
  }
  g.sync();
}

// out = (1 (x) O (x) 1) in, O a 2^mbits square matrix on the middle qubits (any address space)
template <typename T>
QMPS_HDN void bw_apply_mid(const Grp& g, const cx<T>* O, int mbits, int nq, const cx<T>* in, cx<T>* out) {
  const int m = 1 << mbits, mask = (m - 1) << 1, n = 1 << nq;
  for (int idx = g.lane; idx < n; idx += g.size) {
    const int r = (idx >> 1) & (m - 1), base = idx & ~mask;
    const cx<T>* row = O + r * m;
    cx<T> acc = mk<T>(0, 0);
    for (int c = 0; c < m; ++c) cmad(acc, row[c], in[base | (c << 1)]);
    out[idx] = acc;
  }
}

template <typename T> QMPS_HDN T bw_group_sum(const Grp& g, T v, T* red) {
  g.sync();
  red[g.lane] = v;
  g.sync();
  T s = T(0);
  for (int i = 0; i < g.size; ++i) s += red[i];
  return s;
}

// OverlapCalculator.expectation_value (:428-533): mbits = 2 (4x4 operator, 4 qubits) or 4 (16x16, 6 qubits)
template <typename T>
QMPS_HDN T bw_expectation(const Grp& g, const BwWork<T>& W, const cx<T>* O, int mbits) {
  const int cells = mbits == 2 ? 2 : 3, nq = 2 * cells, n = 1 << nq;
  bw_build_state<T>(g, W.k2, W.k1, cells, 0, W.psi, W.tmp);
  bw_apply_mid<T>(g, O, mbits, nq, W.psi, W.tmp);
  g.sync();
  T part = T(0);
  for (int idx = g.lane; idx < n; idx += g.size)
    part += W.psi[idx].re * W.tmp[idx].re + W.psi[idx].im * W.tmp[idx].im;
  return bw_group_sum<T>(g, part, W.red);
}

// ManifoldOverlap.circuit (:228-268) with Mr, Ml in W.mr / W.ml and the 16x16 Wop
template <typename T>
QMPS_HDN cx<T> bw_overlap(const Grp& g, const BwWork<T>& W, const cx<T>* Wop) {
  bw_build_state<T>(g, W.k2, W.k1, 3, 0, W.psi, W.tmp);
  bw_build_state<T>(g, W.b2, W.b1, 3, 1, W.phi, W.tmp);
  bw_apply_mid<T>(g, Wop, 4, 6, W.psi, W.tmp);
  g.sync();
  cx<T> part = mk<T>(0, 0);
  for (int idx = g.lane; idx < 64; idx += g.size) {
    const int a = idx >> 5, b = idx & 1, mid = idx & 0x1E;
    cx<T> coef = mk<T>(0, 0);
    for (int a2 = 0; a2 < 2; ++a2)
      for (int b2 = 0; b2 < 2; ++b2) cmad(coef, W.phi[(a2 << 5) | mid | b2], W.ml[a2 * 2 + a] * W.mr[b2 * 2 + b]);
    cmad(part, coef, W.tmp[idx]);
  }
  const T re = bw_group_sum<T>(g, part.re, W.red);
  const T im = bw_group_sum<T>(g, part.im, W.red);
  return mk<T>(re, im);
}

}  // namespace qmps
