// qmps_b200 ansatz front-end: theta -> U -> A without a circuit simulator library.
//
// The reference builds U(theta) with cirq.unitary(gate) from the gate lists in
// qmps/represent.py:268-423.  Here a gate list is DATA (a GateOp program, built on
// the host by qmps_b200/represent.py) and the kernels interpret it: the columns
// of U that survive unitary_to_tensor (input qubit 0 in |0>, qmps/tools.py:151-154)
// are evolved as D state vectors of nq = log2(D)+1 qubits.
//
// cirq conventions restated: qubit 0 is the most significant bit of
// the row index; rz(t) = exp(-i t Z/2) (rx, ry alike); P**t =
// exp(i pi t/2) (cos(pi t/2) - i sin(pi t/2) P) for P in {X, ZZ, XX, YY}.
#pragma once
#include "core.cuh"

namespace qmps {

enum GateCode : int32_t {
  G_RZ = 0, G_RX = 1, G_RY = 2, G_H = 3, G_CNOT = 4, G_SWAP = 5, G_CZ = 6,
  G_XPOW = 7, G_ZZPOW = 8, G_XXPOW = 9, G_YYPOW = 10, G_X = 11, G_Z = 12,
};

// 32-byte op; angle (or exponent) = scale * theta[param] + offset, param < 0 -> offset only
struct GateOp {
  int32_t code;
  int32_t q0;      // target (1-qubit gates) / control / first qubit
  int32_t q1;      // target of CNOT / second qubit
  int32_t param;
  double scale;
  double offset;
};

template <typename T> QMPS_HD void sincos_t(T x, T* s, T* c);
template <> QMPS_HD void sincos_t<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <> QMPS_HD void sincos_t<float>(float x, float* s, float* c) { sincosf(x, s, c); }

// the angle of op `o` for parameter vector theta with an optional one-coordinate
// shift (rotosolve fan-out: theta + shift * e_coord, qmps/tools.py:432-433)
template <typename T>
QMPS_HD T gate_angle(const GateOp& o, const double* theta, int coord, double shift) {
  double v = o.offset;
  if (o.param >= 0) {
    double th = theta[o.param];
    if (o.param == coord) th += shift;
    v += o.scale * th;
  }
  return (T)v;
}

// (cos, sin) needed by op o: half angle for rotations, pi*t/2 for the ** gates
template <typename T>
QMPS_HD void gate_trig(const GateOp& o, const double* theta, int coord, double shift, T* c, T* s) {
  *c = T(1); *s = T(0);
  switch (o.code) {
    case G_RZ: case G_RX: case G_RY:
      sincos_t<T>(gate_angle<T>(o, theta, coord, shift) * T(0.5), s, c); break;
    case G_XPOW: case G_ZZPOW: case G_XXPOW: case G_YYPOW:
      sincos_t<T>(gate_angle<T>(o, theta, coord, shift) * T(1.5707963267948966), s, c); break;
    default: break;
  }
}

// ---- state in memory (shared), lanes of a group cooperate ---------------------------
// S holds R = 2^nq rows x ncols columns.  With a_layout = 1 the storage IS the MPS
// tensor: element (row = 2i+s, col = j) lives at A[(s*D + i)*D + j]  (D = R/2,
// ncols = D); otherwise plain row-major R x ncols (full unitary, ncols = R).
struct StateLayout {
  int R, ncols, a_layout;
  QMPS_HD int at(int row, int col) const {
    if (a_layout) { int D = R >> 1; return ((row & 1) * D + (row >> 1)) * D + col; }
    return row * ncols + col;
  }
};

template <typename T>
QMPS_HDN void ansatz_eval(const Grp& g, const GateOp* ops, int nops, const double* theta,
                          int coord, double shift, int nq, StateLayout L, cx<T>* S, T* trig) {
  const int R = L.R, nc = L.ncols;
  for (int e = g.lane; e < R * nc; e += g.size) {
    int row = e / nc, col = e - row * nc;
    S[L.at(row, col)] = mk<T>(row == col ? T(1) : T(0), T(0));
  }
  for (int k = g.lane; k < nops; k += g.size) gate_trig<T>(ops[k], theta, coord, shift, &trig[2 * k], &trig[2 * k + 1]);
  g.sync();
  for (int k = 0; k < nops; ++k) {
    const GateOp o = ops[k];
    const T c = trig[2 * k], s = trig[2 * k + 1];
    const int b0 = 1 << (nq - 1 - o.q0);
    const int b1 = 1 << (nq - 1 - o.q1);
    switch (o.code) {
      case G_RZ: case G_Z: case G_CZ: case G_ZZPOW: {           // diagonal gates
        for (int e = g.lane; e < R * nc; e += g.size) {
          int row = e / nc, col = e - row * nc;
          cx<T> ph = mk<T>(1, 0);
          if (o.code == G_RZ) ph = mk<T>(c, (row & b0) ? s : -s);
          else if (o.code == G_Z) ph = mk<T>((row & b0) ? T(-1) : T(1), 0);
          else if (o.code == G_CZ) ph = mk<T>(((row & b0) && (row & b1)) ? T(-1) : T(1), 0);
          else if (((row & b0) != 0) != ((row & b1) != 0)) ph = mk<T>(c * c - s * s, T(2) * c * s);  // e^{i pi t}
          int a = L.at(row, col);
          S[a] = S[a] * ph;
        }
      } break;
      case G_RX: case G_RY: case G_H: case G_XPOW: case G_X: {  // one-qubit mixing gates
        cx<T> m00, m01, m10, m11;
        if (o.code == G_RX) { m00 = m11 = mk<T>(c, 0); m01 = m10 = mk<T>(0, -s); }
        else if (o.code == G_RY) { m00 = m11 = mk<T>(c, 0); m01 = mk<T>(-s, 0); m10 = mk<T>(s, 0); }
        else if (o.code == G_H) { T h = T(0.70710678118654752440); m00 = m01 = m10 = mk<T>(h, 0); m11 = mk<T>(-h, 0); }
        else if (o.code == G_X) { m00 = m11 = mk<T>(0, 0); m01 = m10 = mk<T>(1, 0); }
        else { cx<T> ph = mk<T>(c, s); m00 = m11 = ph * mk<T>(c, 0); m01 = m10 = ph * mk<T>(0, -s); }
        const int half = R >> 1;
        for (int e = g.lane; e < half * nc; e += g.size) {
          int pr = e / nc, col = e - pr * nc;
          int r0 = ((pr & ~(b0 - 1)) << 1) | (pr & (b0 - 1));   // insert a 0 at bit b0
          int a0 = L.at(r0, col), a1 = L.at(r0 | b0, col);
          cx<T> x0 = S[a0], x1 = S[a1];
          S[a0] = m00 * x0 + m01 * x1;
          S[a1] = m10 * x0 + m11 * x1;
        }
      } break;
      case G_CNOT: case G_SWAP: {                                 // permutations
        for (int e = g.lane; e < R * nc; e += g.size) {
          int row = e / nc, col = e - row * nc;
          int partner;
          bool act;
          if (o.code == G_CNOT) { act = (row & b0) && !(row & b1); partner = row | b1; }
          else { act = (row & b0) && !(row & b1); partner = (row & ~b0) | b1; }
          if (act) {
            int a0 = L.at(row, col), a1 = L.at(partner, col);
            cx<T> t = S[a0]; S[a0] = S[a1]; S[a1] = t;
          }
        }
      } break;
      case G_XXPOW: case G_YYPOW: {                               // e^{i pi t/2}(c - i s PP)
        cx<T> ph = mk<T>(c, s);
        cx<T> dd = ph * mk<T>(c, 0), od = ph * mk<T>(0, -s);
        for (int e = g.lane; e < R * nc; e += g.size) {
          int row = e / nc, col = e - row * nc;
          if (row & b0) continue;                                 // handle each pair once
          int partner = row ^ (b0 | b1);
          // <row| PP |partner>: XX -> +1; YY -> -1 if the two bits of `row` are equal, else +1
          T sg = T(1);
          if (o.code == G_YYPOW && (((row & b0) != 0) == ((row & b1) != 0))) sg = T(-1);
          int a0 = L.at(row, col), a1 = L.at(partner, col);
          cx<T> x0 = S[a0], x1 = S[a1];
          S[a0] = dd * x0 + (od * x1) * sg;
          S[a1] = dd * x1 + (od * x0) * sg;
        }
      } break;
      default: break;
    }
    g.sync();
  }
}

// ---- two-qubit programs entirely in registers (thread per problem, D = 2) --------------
// x[row*NC + col], rows 0..3 (qubit 0 = high bit), NC columns.
template <typename T, int NC, int BIT> QMPS_HD void reg2_1q(cx<T>* x, cx<T> m00, cx<T> m01, cx<T> m10, cx<T> m11) {
#pragma unroll
  for (int pr = 0; pr < 2; ++pr) {
    const int r0 = (BIT == 2) ? pr : 2 * pr;      // BIT=2: qubit 0 pairs (0,2),(1,3); BIT=1: (0,1),(2,3)
#pragma unroll
    for (int col = 0; col < NC; ++col) {
      cx<T> x0 = x[r0 * NC + col], x1 = x[(r0 | BIT) * NC + col];
      x[r0 * NC + col] = m00 * x0 + m01 * x1;
      x[(r0 | BIT) * NC + col] = m10 * x0 + m11 * x1;
    }
  }
}
template <typename T, int NC> QMPS_HD void reg2_swap_rows(cx<T>* x, int ra, int rb) {
#pragma unroll
  for (int col = 0; col < NC; ++col) { cx<T> t = x[ra * NC + col]; x[ra * NC + col] = x[rb * NC + col]; x[rb * NC + col] = t; }
}

template <typename T, int NC>
QMPS_HDN void ansatz_reg2(const GateOp* ops, int nops, const double* theta, int coord, double shift, cx<T>* x) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int col = 0; col < NC; ++col) x[r * NC + col] = mk<T>(r == col ? T(1) : T(0), T(0));
  for (int k = 0; k < nops; ++k) {
    const GateOp o = ops[k];
    T c, s;
    gate_trig<T>(o, theta, coord, shift, &c, &s);
    const bool hi = (o.q0 == 0);                   // acts on the high bit (value 2)
    switch (o.code) {
      case G_RZ: {
        cx<T> lo_ph = mk<T>(c, -s), hi_ph = mk<T>(c, s);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool set = hi ? (r & 2) : (r & 1);
#pragma unroll
          for (int col = 0; col < NC; ++col) x[r * NC + col] = x[r * NC + col] * (set ? hi_ph : lo_ph);
        }
      } break;
      case G_Z: {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const bool set = hi ? (r & 2) : (r & 1);
#pragma unroll
          for (int col = 0; col < NC; ++col) if (set) x[r * NC + col] = -x[r * NC + col];
        }
      } break;
      case G_CZ: {
#pragma unroll
        for (int col = 0; col < NC; ++col) x[3 * NC + col] = -x[3 * NC + col];
      } break;
      case G_ZZPOW: {
        cx<T> ph = mk<T>(c * c - s * s, T(2) * c * s);
#pragma unroll
        for (int col = 0; col < NC; ++col) { x[1 * NC + col] = x[1 * NC + col] * ph; x[2 * NC + col] = x[2 * NC + col] * ph; }
      } break;
      case G_RX: case G_RY: case G_H: case G_XPOW: case G_X: {
        cx<T> m00, m01, m10, m11;
        if (o.code == G_RX) { m00 = m11 = mk<T>(c, 0); m01 = m10 = mk<T>(0, -s); }
        else if (o.code == G_RY) { m00 = m11 = mk<T>(c, 0); m01 = mk<T>(-s, 0); m10 = mk<T>(s, 0); }
        else if (o.code == G_H) { T h = T(0.70710678118654752440); m00 = m01 = m10 = mk<T>(h, 0); m11 = mk<T>(-h, 0); }
        else if (o.code == G_X) { m00 = m11 = mk<T>(0, 0); m01 = m10 = mk<T>(1, 0); }
        else { cx<T> ph = mk<T>(c, s); m00 = m11 = ph * mk<T>(c, 0); m01 = m10 = ph * mk<T>(0, -s); }
        if (hi) reg2_1q<T, NC, 2>(x, m00, m01, m10, m11); else reg2_1q<T, NC, 1>(x, m00, m01, m10, m11);
      } break;
      case G_CNOT: {
        if (hi) reg2_swap_rows<T, NC>(x, 2, 3);    // control = qubit 0: |10> <-> |11>
        else reg2_swap_rows<T, NC>(x, 1, 3);       // control = qubit 1: |01> <-> |11>
      } break;
      case G_SWAP: reg2_swap_rows<T, NC>(x, 1, 2); break;
      case G_XXPOW: case G_YYPOW: {
        cx<T> ph = mk<T>(c, s);
        cx<T> dd = ph * mk<T>(c, 0), od = ph * mk<T>(0, -s);
        const T sg = (o.code == G_YYPOW) ? T(-1) : T(1);       // rows 0,3 have equal bits
#pragma unroll
        for (int col = 0; col < NC; ++col) {
          cx<T> a = x[0 * NC + col], b = x[3 * NC + col];
          x[0 * NC + col] = dd * a + (od * b) * sg;
          x[3 * NC + col] = dd * b + (od * a) * sg;
          cx<T> e = x[1 * NC + col], f = x[2 * NC + col];
          x[1 * NC + col] = dd * e + od * f;
          x[2 * NC + col] = dd * f + od * e;
        }
      } break;
      default: break;
    }
  }
}

}  // namespace qmps
