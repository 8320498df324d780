// TEST INFRASTRUCTURE ONLY: compiles the product's device algorithms
// (qmps_b200/csrc/*.cuh, which are __host__ __device__ templates) as plain C++ with a
// cooperating group of ONE lane, so their arithmetic can be checked against the
// oracle on a machine without a GPU.  This library is never loaded by qmps_b200.
#include <stdint.h>
#include <stdlib.h>
#include <vector>

#include "../../qmps_b200/csrc/core.cuh"
#include "../../qmps_b200/csrc/ansatz.cuh"
#include "../../qmps_b200/csrc/generic.cuh"
#include "../../qmps_b200/csrc/d2.cuh"
#include "../../qmps_b200/csrc/envreal.cuh"
#include "../../qmps_b200/csrc/canon.cuh"
#include "../../qmps_b200/csrc/brickwall.cuh"
#include "../../qmps_b200/csrc/fp_d2.cuh"
#include "../../qmps_b200/csrc/tdvp.cuh"

using namespace qmps;
typedef cx<double> zc;

static Grp solo() { Grp g; g.lane = 0; g.size = 1; g.mask = 1; g.cta = 0; return g; }

// real-form environment solve (envreal.cuh): the device kernel's row builder, then the same
// Gauss-Jordan with implicit partial pivoting the kernel runs with one thread per row.
template <int D> static int env_real_one(const zc* A, int d, zc* r) {
  constexpr int n = D * D;
  std::vector<zc> Ap((size_t)d * D * (D + 1));
  for (int q = 0; q < d * n; ++q) Ap[(q / D) * (D + 1) + q % D] = A[q];
  std::vector<double> M((size_t)n * (n + 1));
  for (int e = 0; e < n; ++e) {      // same dispatch as env_real_kernel
    if (d == 2 && (e & 1)) herm_row_cached<double, D, 2>(Ap.data(), D + 1, e, &M[(size_t)e * (n + 1)]);
    else if (d == 2) herm_row_lhoist<double, D, 2>(Ap.data(), D + 1, e, &M[(size_t)e * (n + 1)]);
    else herm_row<double, D>(Ap.data(), D + 1, d, e, &M[(size_t)e * (n + 1)]);
  }
  std::vector<int> done(n, 0), col_of(n, 0);
  std::vector<double> piv(n, 1.0), x(n);
  int bad = 0;
  for (int k = 0; k < n; ++k) {
    int who = -1; double best = -1;
    for (int e = 0; e < n; ++e) if (!done[e] && fabs(M[(size_t)e * (n + 1) + k]) > best) { best = fabs(M[(size_t)e * (n + 1) + k]); who = e; }
    if (!(best > 1e-13)) bad = 1;
    const double* prow = &M[(size_t)who * (n + 1)];
    const double pv = prow[k];
    for (int e = 0; e < n; ++e) {
      if (e == who) continue;
      double f = M[(size_t)e * (n + 1) + k] * (1.0 / pv);
      for (int j = k + 1; j <= n; ++j) M[(size_t)e * (n + 1) + j] -= f * prow[j];
    }
    done[who] = 1; col_of[who] = k; piv[who] = pv;
  }
  for (int e = 0; e < n; ++e) x[col_of[e]] = M[(size_t)e * (n + 1) + n] / piv[e];
  for (int e = 0; e < n; ++e) herm_scatter<double, D>(x.data(), e, r);
  return bad;
}


extern "C" {

int emu_env_d2(int64_t N, const double* a, double* r, double* eta, double* C, int32_t* status) {
  for (int64_t n = 0; n < N; ++n) {
    double e;
    status[n] = env_d2_solve<double, true>((const zc*)a + n * 8, (zc*)r + n * 4, &e, (zc*)C + n * 4);
    eta[2 * n] = e; eta[2 * n + 1] = 0;
  }
  return 0;
}

int emu_env_generic(int d, int D, int64_t N, const double* A, int assume_lc, double* eta, double* r, double* C,
                    int32_t* status) {
  const int n = D * D, ld = n + 1;
  std::vector<zc> E((size_t)n * ld), x(n), Cc(n), w(n), vv(n), rc(n), rs(n);
  std::vector<double> rn(n), red(1);
  std::vector<int> step(n), done(n);
  Grp g = solo();
  for (int64_t p = 0; p < N; ++p) {
    const zc* a = (const zc*)A + p * (size_t)d * n;
    int st;
    zc lam = mk<double>(1, 0);
    if (assume_lc) {
      double er;
      st = env_solve_direct<double>(g, a, d, D, E.data(), ld, x.data(), step.data(), done.data(), red.data(), &er);
      lam = mk<double>(er, 0);
    } else {
      st = leading_eigenpair<double>(g, a, a, d, D, 0, 1, E.data(), ld, w.data(), vv.data(), rc.data(), rs.data(),
                                     rn.data(), x.data(), step.data(), done.data(), &lam);
      hermitise<double>(g, x.data(), D);
      double tr = 0;
      for (int i = 0; i < D; ++i) tr += x[i * D + i].re;
      for (int e = 0; e < n; ++e) x[e] = x[e] * (1.0 / tr);
    }
    int bad = cholesky_lower<double>(g, x.data(), D, Cc.data(), D, D);
    if (bad && st == ST_OK) st = ST_NOT_PD;
    status[p] = st;
    eta[2 * p] = lam.re; eta[2 * p + 1] = lam.im;
    for (int e = 0; e < n; ++e) { ((zc*)r)[p * n + e] = x[e]; ((zc*)C)[p * n + e] = Cc[e]; }
  }
  return 0;
}

int emu_fixed_point(int d, int D, int64_t N, const double* A, const double* B, int left, double* eta, double* vec,
                    int32_t* status) {
  const int n = D * D, ld = n + 1;
  std::vector<zc> H((size_t)n * ld), x(n), w(n), vv(n), rc(n), rs(n);
  std::vector<double> rn(n);
  std::vector<int> step(n), done(n);
  Grp g = solo();
  for (int64_t p = 0; p < N; ++p) {
    zc lam;
    status[p] = leading_eigenpair<double>(g, (const zc*)A + p * (size_t)d * n, (const zc*)B + p * (size_t)d * n, d, D,
                                          left, 1, H.data(), ld, w.data(), vv.data(), rc.data(), rs.data(), rn.data(),
                                          x.data(), step.data(), done.data(), &lam);
    eta[2 * p] = lam.re; eta[2 * p + 1] = lam.im;
    for (int e = 0; e < n; ++e) ((zc*)vec)[p * n + e] = x[e];
  }
  return 0;
}

// all eigenvalues of a dense n x n matrix (row-major, interleaved complex)
int emu_eigvals(int n, const double* M, double* w_out) {
  const int ld = n + 1;
  std::vector<zc> H((size_t)n * ld), w(n), vv(n), rc(n), rs(n);
  std::vector<double> rn(n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) H[i * ld + j] = ((const zc*)M)[i * n + j];
  Grp g = solo();
  hessenberg<double>(g, H.data(), ld, n, vv.data());
  int fail = hqr_eigenvalues<double>(g, H.data(), ld, n, w.data(), rc.data(), rs.data(), rn.data());
  for (int i = 0; i < n; ++i) { w_out[2 * i] = w[i].re; w_out[2 * i + 1] = w[i].im; }
  return fail;
}

int emu_ansatz(const GateOp* ops, int nops, int nq, int64_t N, int P, const double* theta, int full, int coord,
               double shift, double* out) {
  const int R = 1 << nq, nc = full ? R : R / 2;
  std::vector<double> trig(2 * (nops > 0 ? nops : 1));
  StateLayout SL; SL.R = R; SL.ncols = nc; SL.a_layout = full ? 0 : 1;
  Grp g = solo();
  for (int64_t p = 0; p < N; ++p)
    ansatz_eval<double>(g, ops, nops, theta + p * P, coord, shift, nq, SL, (zc*)out + p * (size_t)R * nc, trig.data());
  return 0;
}

// two-qubit register path: out = A[2][2][2] per problem
int emu_ansatz_reg2(const GateOp* ops, int nops, int64_t N, int P, const double* theta, int coord, double shift,
                    double* out) {
  for (int64_t p = 0; p < N; ++p) {
    zc x[8];
    ansatz_reg2<double, 2>(ops, nops, theta + p * P, coord, shift, x);
    zc* a = (zc*)out + p * 8;
    for (int row = 0; row < 4; ++row) for (int j = 0; j < 2; ++j) a[(row & 1) * 4 + (row >> 1) * 2 + j] = x[row * 2 + j];
  }
  return 0;
}

int emu_energy_d2(int64_t N, const double* a, const double* hmat, double* energy) {
  for (int64_t n = 0; n < N; ++n) {
    zc r[4], C[4];
    double eta;
    env_d2_solve<double, true>((const zc*)a + n * 8, r, &eta, C);
    energy[n] = energy_d2<double>((const zc*)a + n * 8, r, (const zc*)hmat);
  }
  return 0;
}

// in: A [N][2][D][D] (two_site = 0) or M [N][4][D][D] (two_site = 1)
int emu_energy_generic(int D, int64_t N, const double* in, int two_site, const double* hmat, double* energy) {
  const int n = D * D, ld = n + 1, d = two_site ? 4 : 2;
  std::vector<zc> E((size_t)n * ld), x(n), tmp(8 * n);
  std::vector<double> red(1);
  std::vector<int> step(n), done(n);
  Grp g = solo();
  for (int64_t p = 0; p < N; ++p) {
    const zc* a = (const zc*)in + p * (size_t)d * n;
    double er;
    env_solve_direct<double>(g, a, d, D, E.data(), ld, x.data(), step.data(), done.data(), red.data(), &er);
    const zc* M = a;
    if (!two_site) { merge_block<double>(g, a, a, 2, 2, D, tmp.data()); M = tmp.data(); }
    energy[p] = energy_from_block<double>(g, M, x.data(), D, (const zc*)hmat, tmp.data() + 4 * n, red.data());
  }
  return 0;
}

int emu_env_real(int d, int D, int64_t N, const double* A, double* r, int32_t* status) {
  for (int64_t p = 0; p < N; ++p) {
    const zc* a = (const zc*)A + p * (size_t)d * D * D;
    zc* out = (zc*)r + p * (size_t)D * D;
    if (D == 2) status[p] = env_real_one<2>(a, d, out);
    else if (D == 4) status[p] = env_real_one<4>(a, d, out);
    else if (D == 8) status[p] = env_real_one<8>(a, d, out);
    else return -1;
  }
  return 0;
}

// canon.cuh: gauge transform (x_kind 0: X = l -> L A L^-1 * scale; 1: X = C -> C^-1 A C)
int emu_gauge(int d, int D, int64_t N, const double* A, const double* X, int x_kind, const double* scale, double* out,
              double* gout, int32_t* status) {
  const int n = D * D;
  std::vector<zc> sA((size_t)d * n), sG(n), sGi(n), sT(n);
  Grp g = solo();
  for (int64_t p = 0; p < N; ++p)
    status[p] = gauge_problem<double>(g, (const zc*)A + p * (size_t)d * n, (const zc*)X + p * (size_t)n, x_kind,
                                      scale ? scale[p] : 1.0, d, D, sA.data(), sG.data(), sGi.data(), sT.data(),
                                      (zc*)out + p * (size_t)d * n, gout ? (zc*)gout + p * (size_t)n : nullptr);
  return 0;
}

// canon.cuh: single-site expectation values; lvec / eta may be null (left-canonical A, tr r = 1)
int emu_expect(int d, int D, int64_t N, const double* A, const double* r, const double* lvec, const double* eta,
               int nops, const double* ops, double* out) {
  const int n = D * D;
  std::vector<zc> sA((size_t)d * n), sR(n), sL(n), sP((size_t)d * n), sQ((size_t)d * d * D), sM(d * d);
  Grp g = solo();
  for (int64_t p = 0; p < N; ++p)
    expect_problem<double>(g, (const zc*)A + p * (size_t)d * n, (const zc*)r + p * (size_t)n,
                           lvec ? (const zc*)lvec + p * (size_t)n : nullptr, eta ? (const zc*)eta + p : nullptr,
                           (const zc*)ops, nops, d, D, sA.data(), sR.data(), sL.data(), sP.data(), sQ.data(), sM.data(),
                           (zc*)out + p * (size_t)nops);
  return 0;
}

// brickwall.cuh: the body of bw_kernel for one problem at a time (mode numbering of kernels_bw.cuh:
// 0 env, 1 apply, 2 expect, 3 overlap, 4 cost).  Arrays as in include/qmps_b200.h; counts 1 broadcast.
int emu_brickwall(int mode, int side, int bra_undaggered, int mbits, int64_t N, int64_t NK, const double* U1,
                  const double* U2, int64_t NB, const double* B1, const double* B2, int64_t NM, const double* Mr,
                  const double* Ml, int64_t NW, const double* Wop, double* mat, double* eta, double* vec,
                  double* overlap, double* real_out, int32_t* status) {
  std::vector<unsigned char> buf(bw_work_bytes<double>(1) + 64);
  Grp g = solo();
  const BwWork<double> W = bw_carve<double>(buf.data(), 1);
  for (int64_t p = 0; p < N; ++p) {
    const zc* u1 = (const zc*)U1 + (NK == 1 ? 0 : p) * 16;
    const zc* u2 = (const zc*)U2 + (NK == 1 ? 0 : p) * 16;
    const zc* b1 = B1 ? (const zc*)B1 + (NB == 1 ? 0 : p) * 16 : nullptr;
    const zc* b2 = B2 ? (const zc*)B2 + (NB == 1 ? 0 : p) * 16 : nullptr;
    const int wsz = (mode == 2 && mbits == 2) ? 16 : 256;
    const zc* wop = Wop ? (const zc*)Wop + (NW == 1 ? 0 : p) * wsz : nullptr;
    bw_load<double>(g, u1, u2, b1, b2, bra_undaggered, W);
    int st = 0;
    if (mode == 0) {
      if (mat) { bw_env_matrix<double>(g, W, side, W.E, 5); for (int e = 0; e < 16; ++e) ((zc*)mat)[p * 16 + e] = W.E[(e >> 2) * 5 + (e & 3)]; }
      zc lam;
      st = bw_exact_environment<double>(g, W, side, &lam, vec != nullptr);
      if (eta) ((zc*)eta)[p] = lam;
      if (vec) for (int e = 0; e < 4; ++e) ((zc*)vec)[p * 4 + e] = W.x[e];
    } else if (mode == 1) {
      bw_env_apply<double>(g, W, (const zc*)Mr + (NM == 1 ? 0 : p) * 4, (zc*)vec + p * 4);
    } else if (mode == 2) {
      real_out[p] = bw_expectation<double>(g, W, wop, mbits);
    } else {
      if (mode == 4) {
        zc lam;
        st = bw_exact_environment<double>(g, W, 0, &lam, 1);
        for (int e = 0; e < 4; ++e) { W.mr[e] = W.x[e]; W.ml[e] = conj(W.x[(e & 1) * 2 + (e >> 1)]); }
        if (eta) ((zc*)eta)[p] = lam;
        if (vec) for (int e = 0; e < 4; ++e) ((zc*)vec)[p * 4 + e] = W.x[e];
      } else {
        for (int e = 0; e < 4; ++e) { W.mr[e] = ((const zc*)Mr)[(NM == 1 ? 0 : p) * 4 + e]; W.ml[e] = ((const zc*)Ml)[(NM == 1 ? 0 : p) * 4 + e]; }
      }
      const zc ov = bw_overlap<double>(g, W, wop);
      if (overlap) ((zc*)overlap)[p] = ov;
      if (real_out) real_out[p] = -(ov.re * ov.re + ov.im * ov.im);
    }
    if (status) status[p] = st;
  }
  return 0;
}

// fp_d2.cuh: register-resident leading eigenvalue of the D = 2 mixed map (the thread body of fp_d2_kernel)
int emu_fp_d2(int d, int64_t N, const double* A, const double* B, int left, double* eta, int32_t* status, double* vec) {
  for (int64_t p = 0; p < N; ++p) {
    zc lam;
    status[p] = fpd2_leading<double>((const zc*)A + p * (size_t)d * 4, (const zc*)B + p * (size_t)d * 4, d, left, &lam,
                                     vec ? (zc*)vec + p * 4 : nullptr);
    eta[2 * p] = lam.re; eta[2 * p + 1] = lam.im;
  }
  return 0;
}

// brickwall.cuh::bw_cost_thread: the thread body of bw_cost_thread_kernel (one ket state, one W); chi is
// built with the same group routines bw_chi_kernel runs
int emu_bw_cost_thread(int64_t N, const double* U1, const double* U2, const double* V1, const double* V2,
                       const double* Wop, double* cost, double* overlap, double* eta, double* Mr, int32_t* status) {
  Grp g = solo();
  zc k2[4], psi[64], tmp[64];
  for (int q = 0; q < 4; ++q) k2[q] = ((const zc*)U2)[q * 4];
  bw_build_state<double>(g, k2, (const zc*)U1, 3, 0, psi, tmp);
  bw_apply_mid<double>(g, (const zc*)Wop, 4, 6, psi, tmp);
  for (int64_t p = 0; p < N; ++p) {
    zc ov, lam, mr[4];
    status[p] = bw_cost_thread<double>((const zc*)U1, k2, tmp, (const zc*)V1 + p * 16, (const zc*)V2 + p * 16, &ov, &lam, mr);
    cost[p] = -(ov.re * ov.re + ov.im * ov.im);
    ((zc*)overlap)[p] = ov; ((zc*)eta)[p] = lam;
    for (int i = 0; i < 4; ++i) ((zc*)Mr)[p * 4 + i] = mr[i];
  }
  return 0;
}

// brickwall.cuh::bw_env_thread: the thread body of bw_env_thread_kernel
int emu_bw_env_thread(int side, int bra_undaggered, int64_t N, const double* U1, const double* U2, const double* B1,
                      const double* B2, double* mat, double* eta, double* vec, int32_t* status) {
  for (int64_t p = 0; p < N; ++p)
    status[p] = bw_env_thread<double>((const zc*)U1 + p * 16, (const zc*)U2 + p * 16, (const zc*)B1 + p * 16,
                                      (const zc*)B2 + p * 16, bra_undaggered, side, (zc*)mat + p * 16, (zc*)eta + p,
                                      (zc*)vec + p * 4);
  return 0;
}

// TDVP tangent vector of left-canonical tensors (tdvp.cuh) and the gauge-back transform
int emu_tdvp_tangent(int d, int D, int64_t N, const double* a, const double* h, int imaginary, double* out, double* energy,
                     int32_t* status) {
  const Grp g = solo();
  const int n = D * D;
  std::vector<zc> E((size_t)n * (n + 1)), x(n), r(n), K(n), rinv(n), Cc(n), Ci(n), Hl(n), AA((size_t)d * d * n), C((size_t)d * d * n),
      G((size_t)d * n);
  std::vector<int> step(n), done(n);
  std::vector<double> red(4);
  for (int64_t k = 0; k < N; ++k) {
    double en = 0;
    status[k] = tdvp_tangent_problem<double>(g, (const zc*)a + k * (size_t)(d * n), (const zc*)h, imaginary, d, D, E.data(), x.data(),
                                             r.data(), K.data(), rinv.data(), Cc.data(), Ci.data(), Hl.data(), AA.data(), C.data(),
                                             G.data(), step.data(), done.data(), red.data(), (zc*)out + k * (size_t)(d * n), &en);
    energy[k] = en;
  }
  return 0;
}

int emu_gauge_back(int d, int D, int64_t N, const double* b, const double* x, const double* scale, double* out) {
  const Grp g = solo();
  const int n = D * D;
  std::vector<zc> sB((size_t)d * n), sX(n), sXi(n), sT(n);
  for (int64_t k = 0; k < N; ++k)
    gauge_back_problem<double>(g, (const zc*)b + k * (size_t)(d * n), (const zc*)x + k * (size_t)n, scale[k], d, D, sB.data(), sX.data(),
                               sXi.data(), sT.data(), (zc*)out + k * (size_t)(d * n));
  return 0;
}

}  // extern "C"
