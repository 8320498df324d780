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

// ---- thread-per-candidate form of Evolve.exact_cost_function (:777-790) ---------------------------------
// One ket state (U1, U2) and one W for a whole population of candidates (V1, V2) -- the TDVP optimiser's
// use: chi = (1 (x) W (x) 1)|psi> is computed once per launch (64 amplitudes, shared), and the candidate's
// work runs in registers: P = V1^dagger U1, the 4 x 4 right map, its eigenvalues (fp_d2.cuh), numpy's
// complex-argmax selection, one inverse iteration in zgeev's gauge -> Mr, then the overlap contracted
// left to right along the bra's matrix-product structure
//   phi[q0,(q1q2),(q3q4),q5] = sum w2[q0,y1] U1_[(y1y2),(q1q2)] w2[y2,y3] U1_[(y3y4),(q3q4)] w2[y4,q5]
// without ever forming the 64-amplitude bra (~450 complex MACs instead of ~2100 per candidate).
#include "fp_d2.cuh"

namespace qmps {

// U1k: ket U1 [16], u2k: U2[:,0] [4], chi [64] (any address space; shared memory on the device).
// V1, V2: the candidate's UNdaggered unitaries [16] each.  Returns the status of the eigen-solve.
template <typename T>
QMPS_HD int bw_cost_thread(const cx<T>* U1k, const cx<T>* u2k, const cx<T>* chi, const cx<T>* V1, const cx<T>* V2,
                           cx<T>* overlap_out, cx<T>* lambda_out, cx<T>* Mr_out) {
  // V1 is NOT held in registers across the eigen-solve (it would cost 64 of them): each stage reloads
  // the four entries it needs (the candidate's 256 bytes stay in L1)
  cx<T> w2[4], u2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { w2[q] = conj(V2[q * 4]); u2[q] = u2k[q]; }       // w2[2y+e] = U2_[0,(y,e)]
  // right map E[(a,b),(c,e)] = sum_xy P[(b,y),(a,x)] u2[x,c] w2[y,e],  P = V1^dagger U1 (built entry by entry)
  cx<T> E[4][4];
  cx<T> lam = mk<T>(0, 0);
  int status = ST_OK;
  cx<T> mr[4];
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) E[r][c] = mk<T>(0, 0);
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      cx<T> vcol[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) vcol[k] = V1[k * 4 + pr];
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        cx<T> pv = mk<T>(0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) cmad_c(pv, U1k[k * 4 + pc], vcol[k]);          // conj(V1[k][pr]) U1[k][pc]
        const int b = pr >> 1, y = pr & 1, a = pc >> 1, x = pc & 1;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const cx<T> pu = pv * u2[2 * x + c];
#pragma unroll
          for (int ee = 0; ee < 2; ++ee) cmad(E[2 * a + b][2 * c + ee], pu, w2[2 * y + ee]);
        }
      }
    }
    if (pass == 0) {
      cx<T> w[4];
      status = fpd2_eigenvalues<T>(E, w);
      int k = 0;
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        // numpy's complex argmax: lexicographic on (real, imag); the running best is selected by value
        bool better = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j == k) better = w[i].re > w[j].re || (w[i].re == w[j].re && w[i].im > w[j].im);
        if (better) k = i;
      }
      lam = w[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) if (i == k) lam = w[i];
    } else {
      fpd2_inverse_iteration<T>(E, lam, mr);
      fpd2_fix_gauge<T>(mr, 1);
    }
  }
  *lambda_out = lam;
  if (Mr_out) {
#pragma unroll
    for (int i = 0; i < 4; ++i) Mr_out[i] = mr[i];
  }
  // wl[x1,q0] = sum_q0' w2[q0',x1] Ml[q0',q0],  Ml = Mr^dagger: Ml[q0',q0] = conj(Mr[q0,q0'])
  cx<T> wl[4];
#pragma unroll
  for (int x1 = 0; x1 < 2; ++x1)
#pragma unroll
    for (int q0 = 0; q0 < 2; ++q0) {
      cx<T> s = mk<T>(0, 0);
#pragma unroll
      for (int qp = 0; qp < 2; ++qp) cmad_c(s, w2[2 * qp + x1], mr[2 * q0 + qp]);
      wl[2 * x1 + q0] = s;
    }
  cx<T> S2[2][4][2];                                  // [x2][(q3q4)][q5']
#pragma unroll
  for (int x2 = 0; x2 < 2; ++x2)
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int q5 = 0; q5 < 2; ++q5) S2[x2][g][q5] = mk<T>(0, 0);
#pragma unroll 1
  for (int h = 0; h < 4; ++h) {                       // h = (q1q2); rolled: bounds the live range of the chi / V1 loads
    cx<T> uh[4];                                      // U1_[(x1x2),h] = conj(V1[h][(x1x2)])
#pragma unroll
    for (int q = 0; q < 4; ++q) uh[q] = conj(V1[h * 4 + q]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      cx<T> t[2][2], s1[2][2];                        // t[q0][q5'], s1[x1][q5']
#pragma unroll
      for (int q0 = 0; q0 < 2; ++q0) {
        const cx<T> c0 = chi[(q0 << 5) | (h << 3) | (g << 1) | 0], c1 = chi[(q0 << 5) | (h << 3) | (g << 1) | 1];
#pragma unroll
        for (int q5 = 0; q5 < 2; ++q5) t[q0][q5] = mr[2 * q5] * c0 + mr[2 * q5 + 1] * c1;
      }
#pragma unroll
      for (int x1 = 0; x1 < 2; ++x1)
#pragma unroll
        for (int q5 = 0; q5 < 2; ++q5) s1[x1][q5] = wl[2 * x1] * t[0][q5] + wl[2 * x1 + 1] * t[1][q5];
#pragma unroll
      for (int x1 = 0; x1 < 2; ++x1)
#pragma unroll
        for (int x2 = 0; x2 < 2; ++x2)
#pragma unroll
          for (int q5 = 0; q5 < 2; ++q5) cmad(S2[x2][g][q5], uh[2 * x1 + x2], s1[x1][q5]);
    }
  }
  cx<T> S4[2][2];                                     // [x4][q5']
#pragma unroll
  for (int x4 = 0; x4 < 2; ++x4)
#pragma unroll
    for (int q5 = 0; q5 < 2; ++q5) S4[x4][q5] = mk<T>(0, 0);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    cx<T> ug[4];                                      // U1_[(x3x4),g] = conj(V1[g][(x3x4)])
#pragma unroll
    for (int q = 0; q < 4; ++q) ug[q] = conj(V1[g * 4 + q]);
#pragma unroll
    for (int x3 = 0; x3 < 2; ++x3) {
      cx<T> s3[2];                                    // S3[x3][g][q5'] = sum_x2 w2[x2,x3] S2[x2][g][q5']
#pragma unroll
      for (int q5 = 0; q5 < 2; ++q5) s3[q5] = w2[x3] * S2[0][g][q5] + w2[2 + x3] * S2[1][g][q5];
#pragma unroll
      for (int x4 = 0; x4 < 2; ++x4)
#pragma unroll
        for (int q5 = 0; q5 < 2; ++q5) cmad(S4[x4][q5], ug[2 * x3 + x4], s3[q5]);
    }
  }
  cx<T> ov = mk<T>(0, 0);
#pragma unroll
  for (int x4 = 0; x4 < 2; ++x4)
#pragma unroll
    for (int q5 = 0; q5 < 2; ++q5) cmad(ov, w2[2 * x4 + q5], S4[x4][q5]);
  *overlap_out = ov;
  return status;
}

}  // namespace qmps

// ---- thread-per-problem form of exact_environment_circuit + exact_environment (:322-352, :394-426) ------------
namespace qmps {

// U1, U2: ket unitaries [16]; B1, B2: what the reference passes as U1_, U2_ (bra_undaggered = 0) or the candidate
// unitaries themselves (1).  side 0 = right, 1 = left.  mat_out (optional, 16 entries): the 4 x 4 map.
template <typename T>
QMPS_HD int bw_env_thread(const cx<T>* U1, const cx<T>* U2, const cx<T>* B1, const cx<T>* B2, int bra_undaggered,
                          int side, cx<T>* mat_out, cx<T>* lambda_out, cx<T>* vec_out) {
  cx<T> u2[4], w2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { u2[q] = U2[q * 4]; w2[q] = bra_undaggered ? conj(B2[q * 4]) : B2[q]; }
  cx<T> E[4][4];
  cx<T> lam = mk<T>(0, 0);
  int status = ST_OK;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) E[r][c] = mk<T>(0, 0);
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      cx<T> brow[4];                                   // row pr of U1_
#pragma unroll
      for (int k = 0; k < 4; ++k) brow[k] = bra_undaggered ? conj(B1[k * 4 + pr]) : B1[pr * 4 + k];
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        cx<T> pv = mk<T>(0, 0);                        // P[pr][pc] = sum_k U1_[pr][k] U1[k][pc]
#pragma unroll
        for (int k = 0; k < 4; ++k) cmad(pv, brow[k], U1[k * 4 + pc]);
        // right: P[(b,y),(a,x)] u2[x,c] w2[y,e] -> E[(a,b),(c,e)];  left: P[(y,b),(x,a)] u2[c,x] w2[e,y]
        const int r1 = pr >> 1, r0 = pr & 1, c1 = pc >> 1, c0 = pc & 1;
        const int a = side == 0 ? c1 : c0, b = side == 0 ? r1 : r0, x = side == 0 ? c0 : c1, y = side == 0 ? r0 : r1;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const cx<T> pu = pv * (side == 0 ? u2[2 * x + c] : u2[2 * c + x]);
#pragma unroll
          for (int ee = 0; ee < 2; ++ee) cmad(E[2 * a + b][2 * c + ee], pu, side == 0 ? w2[2 * y + ee] : w2[2 * ee + y]);
        }
      }
    }
    if (pass == 0) {
      if (mat_out) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) mat_out[r * 4 + c] = E[r][c];
      }
      cx<T> w[4];
      status = fpd2_eigenvalues<T>(E, w);
      int k = 0;
#pragma unroll
      for (int i = 1; i < 4; ++i) {
        bool better = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j == k) better = w[i].re > w[j].re || (w[i].re == w[j].re && w[i].im > w[j].im);
        if (better) k = i;
      }
      lam = w[0];
#pragma unroll
      for (int i = 1; i < 4; ++i) if (i == k) lam = w[i];
      *lambda_out = lam;
      if (!vec_out) return status;
    } else {
      cx<T> x[4];
      fpd2_inverse_iteration<T>(E, lam, x);
      fpd2_fix_gauge<T>(x, 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) vec_out[i] = x[i];
    }
  }
  return status;
}

}  // namespace qmps
