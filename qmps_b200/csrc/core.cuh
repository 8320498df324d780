// qmps_b200 core: complex arithmetic and small dense linear algebra on a matrix
// held in shared (or global workspace) memory, executed by a GROUP of cooperating
// lanes (a sub-warp slice, a warp, or a whole CTA).
//
// Conventions used by every routine here:
//   * matrices are row-major cx<T> with leading dimension `ld` (padded to an odd
//     number of 16-byte elements so lane-over-rows access is bank-conflict free);
//   * decisions that steer control flow (pivots, shifts, deflation) are computed
//     REDUNDANTLY by every lane from the same memory with the same instruction
//     sequence, so they are bit-identical across the group and the group never
//     diverges around a sync();
//   * every read-after-write across lanes is separated by g.sync().
//
// The same source compiles as plain C++ (QMPS_HOST_EMU, group of one lane) for the
// CPU-side algorithm tests under tests/host_emu/ -- that build is test
// infrastructure, never a product path.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define QMPS_HD __host__ __device__ __forceinline__
#define QMPS_HDN __host__ __device__
#else
#define QMPS_HD inline
#define QMPS_HDN
#endif

namespace qmps {

// ---- status codes (mirrored in include/qmps_b200.h) -------------------------
enum : int32_t {
  ST_OK = 0,
  ST_NOT_PD = 1,        // Cholesky of r failed  -> numpy.linalg.LinAlgError upstream
  ST_NO_CONVERGE = 2,   // QR iteration hit its sweep limit
  ST_SINGULAR = 3,      // fixed-point system singular (degenerate leading eigenvalue)
};

template <typename T> struct eps_of;
template <> struct eps_of<double> { static QMPS_HD double v() { return 2.220446049250313e-16; } };
template <> struct eps_of<float> { static QMPS_HD float v() { return 1.1920929e-7f; } };

// ---- complex numbers -----------------------------------------------------------
template <typename T> struct cx {
  T re, im;
};
template <typename T> QMPS_HD cx<T> mk(T re, T im) { cx<T> z; z.re = re; z.im = im; return z; }
template <typename T> QMPS_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> QMPS_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
template <typename T> QMPS_HD cx<T> operator-(cx<T> a) { return mk<T>(-a.re, -a.im); }
template <typename T> QMPS_HD cx<T> operator*(cx<T> a, cx<T> b) {
  return mk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> QMPS_HD cx<T> operator*(cx<T> a, T s) { return mk<T>(a.re * s, a.im * s); }
template <typename T> QMPS_HD cx<T> conj(cx<T> a) { return mk<T>(a.re, -a.im); }
template <typename T> QMPS_HD T norm2(cx<T> a) { return a.re * a.re + a.im * a.im; }
template <typename T> QMPS_HD T cabs(cx<T> a) { return sqrt(norm2(a)); }
template <typename T> QMPS_HD T cabs1(cx<T> a) { return fabs(a.re) + fabs(a.im); }
// fused multiply-add on either precision (one rounding; DFMA / FFMA on the device)
QMPS_HD double fma_t(double a, double b, double c) { return fma(a, b, c); }
QMPS_HD float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
// acc += a*b          (4 FMAs, no separate multiplies / adds)
template <typename T> QMPS_HD void cmad(cx<T>& acc, cx<T> a, cx<T> b) {
  acc.re = fma_t(-a.im, b.im, fma_t(a.re, b.re, acc.re));
  acc.im = fma_t(a.im, b.re, fma_t(a.re, b.im, acc.im));
}
// acc += a*conj(b)
template <typename T> QMPS_HD void cmad_c(cx<T>& acc, cx<T> a, cx<T> b) {
  acc.re = fma_t(a.im, b.im, fma_t(a.re, b.re, acc.re));
  acc.im = fma_t(-a.re, b.im, fma_t(a.im, b.re, acc.im));
}
// acc -= a*b
template <typename T> QMPS_HD void cmsub(cx<T>& acc, cx<T> a, cx<T> b) {
  acc.re = fma_t(a.im, b.im, fma_t(-a.re, b.re, acc.re));
  acc.im = fma_t(-a.im, b.re, fma_t(-a.re, b.im, acc.im));
}
template <typename T> QMPS_HD cx<T> cinv(cx<T> a) {
  T d = T(1) / norm2(a);
  return mk<T>(a.re * d, -a.im * d);
}
template <typename T> QMPS_HD cx<T> cdiv(cx<T> a, cx<T> b) { return a * cinv(b); }
template <typename T> QMPS_HD cx<T> csqrt(cx<T> z) {
  T m = cabs(z);
  if (m == T(0)) return mk<T>(0, 0);
  T a = sqrt((m + fabs(z.re)) * T(0.5));
  T b = z.im / (a + a);
  if (z.re >= T(0)) return mk<T>(a, b);
  return mk<T>(fabs(b), z.im < T(0) ? -a : a);
}

// 1/sqrt(x): one MUFU + Newton steps on the device instead of a square root AND a division
QMPS_HD double rsqrt_hd(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
QMPS_HD float rsqrt_hd(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}

// ---- the cooperating group -------------------------------------------------------
struct Grp {
  int lane;       // 0 .. size-1
  int size;       // lanes cooperating on one problem
  unsigned mask;  // __syncwarp mask when the group is (part of) one warp
  int cta;        // 1: the group is the whole CTA (__syncthreads)
  QMPS_HD void sync() const {
#if defined(__CUDA_ARCH__)
    if (cta) __syncthreads(); else __syncwarp(mask);
#endif
  }
};

// ---- transfer matrix --------------------------------------------------------------
// E[(i,k),(j,l)] = sum_s A[s,i,j] conj(B[s,k,l])   (SURVEY A.1;
// reference definition new_tdvp/EnvironmentParamSensitivity.py:37-38)
// A, B: [d][D][D] row-major (any address space).  E: n x n, n = D*D, leading dim ld.
template <typename T>
QMPS_HDN void build_transfer(const Grp& g, const cx<T>* A, const cx<T>* B, int d, int D,
                             cx<T>* E, int ld) {
  const int n = D * D;
  for (int e = g.lane; e < n * n; e += g.size) {
    int row = e / n, col = e - row * n;
    int i = row / D, k = row - i * D;
    int j = col / D, l = col - j * D;
    cx<T> acc = mk<T>(0, 0);
    for (int s = 0; s < d; ++s) cmad_c(acc, A[(s * D + i) * D + j], B[(s * D + k) * D + l]);
    E[row * ld + col] = acc;
  }
}

// ---- LU with partial pivoting on an augmented system [M | b] -------------------
// M: n x (n+1) (column n is the right-hand side), solved in place; x[n] receives
// the solution.  Rows are never swapped physically: step_row[k] records the pivot
// row of step k, done[i] marks used rows.  `tiny`: pivots smaller than this are
// either reported (replace_tiny = 0 -> returns 1) or replaced by `tiny`
// (inverse iteration).  Returns 0 on success.
template <typename T>
QMPS_HDN int lu_solve_aug(const Grp& g, cx<T>* M, int ld, int n, cx<T>* x, int* step_row,
                          int* done, T tiny, int replace_tiny) {
  int bad = 0;
  for (int i = g.lane; i < n; i += g.size) done[i] = 0;
  g.sync();
  for (int k = 0; k < n; ++k) {
    // pivot search (redundant on every lane)
    int p = -1;
    T best = T(-1);
    for (int i = 0; i < n; ++i) {
      if (done[i]) continue;
      T a = norm2(M[i * ld + k]);
      if (a > best) { best = a; p = i; }
    }
    if (p < 0) {                     // column is all-NaN: take any unused row, flag it
      bad = 1;
      for (int i = 0; i < n && p < 0; ++i) if (!done[i]) p = i;
    }
    cx<T> pv = M[p * ld + k];
    if (!(best >= tiny * tiny)) {   // tiny (or NaN) pivot
      if (replace_tiny) pv = mk<T>(tiny, 0); else bad = 1;
    }
    cx<T> inv = cinv(pv);
    g.sync();                        // everyone has read column k and done[]
    if (g.lane == 0) { step_row[k] = p; done[p] = 1; M[p * ld + k] = pv; }
    // rank-1 update of the not-yet-used rows, columns k+1 .. n (n = rhs)
    const int w = n - k;             // columns to touch per row
    for (int e = g.lane; e < n * w; e += g.size) {
      int i = e / w, j = k + 1 + (e - i * w);
      if (i == p || done[i]) continue;
      cx<T> f = M[i * ld + k] * inv;
      cmsub(M[i * ld + j], f, M[p * ld + j]);
    }
    g.sync();
  }
  // back substitution on the pivot rows in reverse order
  for (int k = n - 1; k >= 0; --k) {
    int p = step_row[k];
    cx<T> xk = cdiv(M[p * ld + n], M[p * ld + k]);   // row p is not written below
    if (g.lane == 0) x[k] = xk;
    for (int m = g.lane; m < k; m += g.size) {
      int q = step_row[m];
      cmsub(M[q * ld + n], M[q * ld + k], xk);
    }
    g.sync();
  }
  return bad;
}

// ---- reduction to upper Hessenberg form (unblocked Householder) -----------------
// In place on the n x n matrix H; vv[n] is scratch.  Similarity transform, so the
// spectrum is preserved; only eigenvalues are wanted, the reflectors are dropped.
template <typename T>
QMPS_HDN void hessenberg(const Grp& g, cx<T>* H, int ld, int n, cx<T>* vv) {
  for (int k = 0; k + 2 < n; ++k) {
    cx<T> alpha = H[(k + 1) * ld + k];
    T xn2 = T(0);
    for (int i = k + 2; i < n; ++i) xn2 += norm2(H[i * ld + k]);
    if (xn2 == T(0) && alpha.im == T(0)) continue;          // already reduced (uniform decision)
    T beta = sqrt(norm2(alpha) + xn2);
    if (alpha.re > T(0)) beta = -beta;
    cx<T> tau = mk<T>((beta - alpha.re) / beta, -alpha.im / beta);
    cx<T> scal = cinv(alpha - mk<T>(beta, 0));
    g.sync();                                               // all lanes hold alpha/xn2
    for (int i = k + 1 + g.lane; i < n; i += g.size) {
      if (i == k + 1) { vv[i] = mk<T>(1, 0); H[i * ld + k] = mk<T>(beta, 0); }
      else { vv[i] = H[i * ld + k] * scal; H[i * ld + k] = mk<T>(0, 0); }
    }
    g.sync();
    // left:  H[k+1:, k+1:] -= conj(tau) v (v^H H)
    cx<T> ctau = conj(tau);
    for (int j = k + 1 + g.lane; j < n; j += g.size) {
      cx<T> s = mk<T>(0, 0);
      for (int i = k + 1; i < n; ++i) cmad(s, conj(vv[i]), H[i * ld + j]);
      s = s * ctau;
      for (int i = k + 1; i < n; ++i) cmsub(H[i * ld + j], vv[i], s);
    }
    g.sync();
    // right: H[:, k+1:] -= tau (H v) v^H
    for (int i = g.lane; i < n; i += g.size) {
      cx<T> s = mk<T>(0, 0);
      for (int j = k + 1; j < n; ++j) cmad(s, H[i * ld + j], vv[j]);
      s = s * tau;
      for (int j = k + 1; j < n; ++j) cmsub(H[i * ld + j], s, conj(vv[j]));
    }
    g.sync();
  }
}

// ---- complex single-shift QR iteration on a Hessenberg matrix -------------------
// Eigenvalues only (active-window updates), explicit QR sweep: all left Givens
// rotations (one sync each; lanes own columns), then all right rotations with NO
// sync (lanes own rows).  w[n] receives the eigenvalues.  rc/rs/rn: scratch
// arrays of n entries for the rotations.  Returns 0, or 1 if some eigenvalue
// needed more than `maxit` sweeps (w then holds the current diagonal).
template <typename T>
QMPS_HDN int hqr_eigenvalues(const Grp& g, cx<T>* H, int ld, int n, cx<T>* w, cx<T>* rc,
                             cx<T>* rs, T* rn) {
  const T eps = eps_of<T>::v();
  const int maxit = 60;
  int fail = 0;
  int en = n - 1;
  g.sync();
  while (en >= 0) {
    int its = 0;
    for (;;) {
      // -- look for a negligible sub-diagonal entry (redundant scan)
      int l = en;
      for (; l > 0; --l) {
        T s = cabs1(H[(l - 1) * ld + (l - 1)]) + cabs1(H[l * ld + l]);
        if (s == T(0)) s = T(1);
        if (cabs1(H[l * ld + (l - 1)]) <= eps * s) break;
      }
      if (l == en || its >= maxit) {
        if (l != en) fail = 1;
        cx<T> ev = H[en * ld + en];
        g.sync();
        if (g.lane == 0) {
          w[en] = ev;
          if (l > 0) H[l * ld + (l - 1)] = mk<T>(0, 0);
        }
        --en;
        g.sync();
        break;
      }
      // -- shift
      cx<T> sh;
      if (its == 10 || its == 20 || its == 30 || its == 40) {
        T t = fabs(H[en * ld + (en - 1)].re) + (en >= 2 ? fabs(H[(en - 1) * ld + (en - 2)].re) : T(0));
        sh = H[en * ld + en] + mk<T>(t, 0);
      } else {
        cx<T> a = H[(en - 1) * ld + (en - 1)], b = H[(en - 1) * ld + en];
        cx<T> c = H[en * ld + (en - 1)], d = H[en * ld + en];
        sh = d;
        cx<T> bc = b * c;
        if (bc.re != T(0) || bc.im != T(0)) {
          cx<T> y = (a - d) * T(0.5);
          cx<T> z = csqrt(y * y + bc);
          if (y.re * z.re + y.im * z.im < T(0)) z = -z;
          sh = d - cdiv(bc, y + z);
        }
      }
      g.sync();                                  // scan/shift reads done before writes
      if (g.lane == 0 && l > 0) H[l * ld + (l - 1)] = mk<T>(0, 0);
      for (int i = l + g.lane; i <= en; i += g.size) H[i * ld + i] = H[i * ld + i] - sh;
      g.sync();
      // -- left rotations: R = G_en ... G_{l+1} (H - sh)
      for (int i = l + 1; i <= en; ++i) {
        cx<T> f = H[(i - 1) * ld + (i - 1)], gg = H[i * ld + (i - 1)];
        const T nr2 = norm2(f) + norm2(gg);
        T nr = T(0);
        cx<T> c, s;
        if (nr2 == T(0)) { c = mk<T>(1, 0); s = mk<T>(0, 0); }
        else { const T inr = rsqrt_hd(nr2); nr = nr2 * inr; c = f * inr; s = gg * inr; }
        if (g.lane == 0) { rc[i] = c; rs[i] = s; rn[i] = nr; }
        // columns j >= i only: column i-1 is fixed up after the loop (nobody reads
        // rows i-1,i of column i-1 again during the left phase)
        for (int j = i + g.lane; j <= en; j += g.size) {
          cx<T> p = H[(i - 1) * ld + j], q = H[i * ld + j];
          H[(i - 1) * ld + j] = conj(c) * p + conj(s) * q;
          H[i * ld + j] = c * q - s * p;
        }
        g.sync();
      }
      for (int i = l + 1 + g.lane; i <= en; i += g.size) {
        H[(i - 1) * ld + (i - 1)] = mk<T>(rn[i], 0);
        H[i * ld + (i - 1)] = mk<T>(0, 0);
      }
      g.sync();
      // -- right rotations: H' = R G_{l+1}^H ... G_en^H + sh ; lane owns rows
      for (int i = l + g.lane; i <= en; i += g.size) {
        int j0 = (i > l + 1) ? i : l + 1;
        for (int j = j0; j <= en; ++j) {
          cx<T> c = rc[j], s = rs[j];
          cx<T> xx = H[i * ld + (j - 1)], yy = H[i * ld + j];
          H[i * ld + (j - 1)] = xx * c + yy * s;
          H[i * ld + j] = yy * conj(c) - xx * conj(s);
        }
      }
      g.sync();
      for (int i = l + g.lane; i <= en; i += g.size) H[i * ld + i] = H[i * ld + i] + sh;
      g.sync();
      ++its;
    }
  }
  return fail;
}

// ---- branch-free forms for the Wilkinson shift of the QR sweeps ------------------------------------------------
// The shift needs a complex square root and a complex division per sweep: two IEEE square roots and two IEEE divisions,
// each with a slow-path call on the device.  A shift only has to be the SAME number when it is subtracted and added
// back, not a correctly rounded one, so on the device these use MUFU seeds + two Newton steps (an ulp or two) and no
// branches; on the host (tests/host_emu) they are the exact forms above.
QMPS_HD double rsq_nb(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}
QMPS_HD float rsq_nb(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
QMPS_HD double rcp_nb(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(fma(-x, y, 1.0), y, y);
  y = fma(fma(-x, y, 1.0), y, y);
  return y;
#else
  return 1.0 / x;
#endif
}
QMPS_HD float rcp_nb(float x) { return 1.0f / x; }
template <typename T> struct nb_floor;
template <> struct nb_floor<double> { static QMPS_HD double v() { return 1e-280; } };
template <> struct nb_floor<float> { static QMPS_HD float v() { return 1e-34f; } };
// sqrt(z), principal branch (same branch choice as csqrt); 0 for |z|^2 below the floor
template <typename T> QMPS_HD cx<T> csqrt_nb(cx<T> z) {
  const T n2 = norm2(z);
  const bool ok = n2 > nb_floor<T>::v();
  const T m = n2 * rsq_nb(ok ? n2 : T(1));                       // |z|
  const T h = (m + fabs(z.re)) * T(0.5);                         // >= |z| / 2 > 0
  const T ra = rsq_nb(ok ? h : T(1));
  const T a = ok ? h * ra : T(0);                                // sqrt(h)
  const T b = ok ? z.im * (T(0.5) * ra) : T(0);                  // z.im / (2 a)
  cx<T> r;
  r.re = z.re >= T(0) ? a : fabs(b);
  r.im = z.re >= T(0) ? b : (z.im < T(0) ? -a : a);
  return r;
}
// a / b; a itself times zero (i.e. 0) for |b|^2 below the floor -- the caller's shift then stays the corner entry
template <typename T> QMPS_HD cx<T> cdiv_nb(cx<T> a, cx<T> b) {
  const T n2 = norm2(b);
  const bool ok = n2 > nb_floor<T>::v();
  const T d = ok ? rcp_nb(ok ? n2 : T(1)) : T(0);
  return a * mk<T>(b.re * d, -b.im * d);
}

// index of the eigenvalue of largest modulus (first one on ties)
template <typename T> QMPS_HD int argmax_abs(const cx<T>* w, int n) {
  int k = 0;
  T best = norm2(w[0]);
  for (int i = 1; i < n; ++i) {
    T a = norm2(w[i]);
    if (a > best) { best = a; k = i; }
  }
  return k;
}

// ---- Cholesky of a Hermitian D x D matrix (lower factor, positive diagonal) -----
// R: D x D (ld_r) -- only its lower triangle is read.  C: D x D (ld_c) fully
// written (upper part zero).  Returns 0, or 1 if R is not positive definite
// (same criterion as LAPACK zpotrf: a non-positive or NaN pivot).
template <typename T>
QMPS_HDN int cholesky_lower(const Grp& g, const cx<T>* R, int ld_r, cx<T>* C, int ld_c, int D) {
  int bad = 0;
  for (int e = g.lane; e < D * D; e += g.size) {
    int i = e / D, j = e - i * D;
    if (j > i) C[i * ld_c + j] = mk<T>(0, 0);
  }
  g.sync();
  for (int j = 0; j < D; ++j) {
    T dsum = R[j * ld_r + j].re;
    for (int k = 0; k < j; ++k) dsum -= norm2(C[j * ld_c + k]);
    if (!(dsum > T(0))) { bad = 1; dsum = T(1); }          // uniform decision
    T cjj = sqrt(dsum), inv = T(1) / cjj;
    if (g.lane == 0) C[j * ld_c + j] = mk<T>(cjj, 0);
    for (int i = j + 1 + g.lane; i < D; i += g.size) {
      cx<T> s = R[i * ld_r + j];
      for (int k = 0; k < j; ++k) cmad_c(s, -C[i * ld_c + k], C[j * ld_c + k]);
      C[i * ld_c + j] = s * inv;
    }
    g.sync();
  }
  return bad;
}

}  // namespace qmps
