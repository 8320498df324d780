// qmps_b200 PXP scar-dynamics step cost (the reference's scars.py:76-155): another caller of the
// D = 2 mixed fixed point, with a two-site unit cell and a four-site evolution gate.
//
//   A(theta, phi)[0] = [[0, i e^{-i phi}], [0, 0]],  A(theta, phi)[1] = [[cos theta, 0], [sin theta, 0]]     (scars.py:70-73)
//   M  = merge(A(th1, ph1), A(th2, ph2))   for the current parameters [th1, ph1, ph2, th2]   (d = 4, D = 2)
//   M' = the same for the candidate parameters
//   (eta, r) = Map(M, M').right_fixed_point()    -- unit norm, zgeev gauge (component of largest modulus real positive)
//   cost = -2 |<0^8| C |0^8>|, the 8-qubit read-out of scars.py:91-110, which is the contraction
//       Phi[i, (s,t), b] = (M^s M^t)[i, b]   (two unit cells),      T = Phi'^dagger W Phi   (W: 16 x 16)
//       amplitude = 1/2 sum_{x,y,i,b} conj(r[i][x]) r[b][y] T[(x,y),(i,b)]
//   (the embeddings put_env_on_right_site(r^dagger) / put_env_on_left_site(r) contribute exactly r^dagger and r^T on
//   the |0> ancilla blocks, qmps/time_evolve_tools.py:38-70; oracle/scars.py checks this against the gate-by-gate
//   state vector and tests/golden/ref_scars.npz holds the reference's own numbers).
// One warp per candidate; the eigen-solve is the register code of fp_d2.cuh run redundantly by every lane.
#pragma once
#include <cuda_runtime.h>
#include "fp_d2.cuh"

namespace qmps {

// merged two-site tensor of the PXP ansatz: out[s = 2 sa + sb][i][j] = sum_k A1[sa][i][k] A2[sb][k][j]
template <typename T> QMPS_HD cx<T> scars_a(int s, int i, int j, T c, T sn, T cph, T sph) {
  // A[0][0][1] = i e^{-i phi} = sin(phi) + i cos(phi);  A[1][0][0] = cos(theta);  A[1][1][0] = sin(theta)
  if (s == 0) return (i == 0 && j == 1) ? mk<T>(sph, cph) : mk<T>(0, 0);
  return j == 0 ? mk<T>(i == 0 ? c : sn, 0) : mk<T>(0, 0);
}
template <typename T> QMPS_HD cx<T> scars_merged(int e, const double* p) {
  const int s = e >> 2, i = (e >> 1) & 1, j = e & 1, sa = s >> 1, sb = s & 1;
  const T c1 = (T)cos(p[0]), s1 = (T)sin(p[0]), cp1 = (T)cos(p[1]), sp1 = (T)sin(p[1]);
  const T cp2 = (T)cos(p[2]), sp2 = (T)sin(p[2]), c2 = (T)cos(p[3]), s2 = (T)sin(p[3]);
  cx<T> acc = mk<T>(0, 0);
  for (int k = 0; k < 2; ++k) cmad(acc, scars_a<T>(sa, i, k, c1, s1, cp1, sp1), scars_a<T>(sb, k, j, c2, s2, cp2, sp2));
  return acc;
}

#if defined(__CUDACC__)

template <typename T>
__global__ void __launch_bounds__(128)
scars_cost_kernel(int64_t N, const double* __restrict__ params, int64_t NC, const double* __restrict__ current,
                  const cx<T>* __restrict__ W, T* __restrict__ cost, cx<T>* __restrict__ eta, int32_t* __restrict__ status) {
  __shared__ cx<T> sW[256];
  __shared__ cx<T> sbuf[4][16 + 16 + 64 + 64 + 64 + 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < 256; e += blockDim.x) sW[e] = W[e];
  __syncthreads();
  cx<T>* sM = sbuf[warp];
  cx<T>* sMp = sM + 16;
  cx<T>* sPhi = sMp + 16;
  cx<T>* sPhip = sPhi + 64;
  cx<T>* sWP = sPhip + 64;
  cx<T>* sT = sWP + 64;
  for (int64_t cand = (int64_t)blockIdx.x * 4 + warp; cand < N; cand += (int64_t)gridDim.x * 4) {
    const double* pc = current + (NC > 1 ? cand : 0) * 4;
    const double* pp = params + cand * 4;
    if (lane < 16) sM[lane] = scars_merged<T>(lane, pc); else sMp[lane - 16] = scars_merged<T>(lane - 16, pp);
    __syncwarp();
    // leading eigenpair of E[(i,k),(j,l)] = sum_s M[s,i,j] conj(M'[s,k,l])
    cx<T> E[4][4], lam, x[4];
    fpd2_build<T>(sM, sMp, 4, 0, E);
    const int st = fpd2_leading_of<T>(E, &lam);
    fpd2_build<T>(sM, sMp, 4, 0, E);
    fpd2_inverse_iteration<T>(E, lam, x);
    fpd2_fix_gauge<T>(x, 1);                                   // r[j][l] = x[2 j + l], unit norm
    // two unit cells: Phi[i][p = 4 s + t][b] at index i * 32 + p * 2 + b
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = lane + 32 * h, i = idx >> 5, p = (idx >> 1) & 15, b = idx & 1, s = p >> 2, t = p & 3;
      cx<T> a = mk<T>(0, 0), ap = mk<T>(0, 0);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        cmad(a, sM[s * 4 + i * 2 + k], sM[t * 4 + k * 2 + b]);
        cmad(ap, sMp[s * 4 + i * 2 + k], sMp[t * 4 + k * 2 + b]);
      }
      sPhi[idx] = a; sPhip[idx] = ap;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                // WP[i][q][b] = sum_p W[q][p] Phi[i][p][b]
      const int idx = lane + 32 * h, i = idx >> 5, q = (idx >> 1) & 15, b = idx & 1;
      cx<T> a = mk<T>(0, 0);
      for (int p = 0; p < 16; ++p) cmad(a, sW[q * 16 + p], sPhi[i * 32 + p * 2 + b]);
      sWP[idx] = a;
    }
    __syncwarp();
    cx<T> term = mk<T>(0, 0);
    if (lane < 16) {                                             // T[x][y][i][b] and its weight conj(r[i][x]) r[b][y]
      const int xx = lane >> 3, yy = (lane >> 2) & 1, i = (lane >> 1) & 1, b = lane & 1;
      cx<T> a = mk<T>(0, 0);
      for (int q = 0; q < 16; ++q) cmad(a, conj(sPhip[xx * 32 + q * 2 + yy]), sWP[i * 32 + q * 2 + b]);
      term = conj(x[2 * i + xx]) * x[2 * b + yy] * a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      term.re += __shfl_xor_sync(0xffffffffu, term.re, o);
      term.im += __shfl_xor_sync(0xffffffffu, term.im, o);
    }
    if (lane == 0) {
      cost[cand] = -cabs(term);                                  // -2 |amplitude|, amplitude = term / 2
      if (eta) eta[cand] = lam;
      if (status) status[cand] = st;
    }
    __syncwarp();
  }
}

#endif

}  // namespace qmps
