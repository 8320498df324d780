/* qmps_b200 -- C ABI of the B200-native classical inner loop of qmps.
 *
 * The reference (fergusfinn/qmps) is pure Python and has no FFI; its boundary for
 * this path is the set of Python callables of SURVEY.md section 8(b).  Every entry
 * point below names the reference callable (file:line under the reference root) it
 * replaces.  INTEGRATION.md shows the ctypes stubs a maintainer adds to qmps/tools.py
 * etc. to route those callables here.
 *
 * Conventions
 *   - complex numbers are interleaved (re, im) pairs: complex128 = 2 x double
 *     (numpy.complex128 / cuDoubleComplex layout), complex64 = 2 x float.
 *   - `dtype`: QMPS_C128 or QMPS_C64 selects the arithmetic AND the element type of
 *     every complex/real array argument (theta and shift arrays are always double).
 *   - all arrays are contiguous, batch-major, row-major; tensors A are [d][D][D]
 *     with A[s][i][j] as in unitary_to_tensor (qmps/tools.py:151-154).
 *   - functions without the _host suffix take DEVICE pointers and enqueue work on
 *     `stream` (a cudaStream_t passed as void*; NULL = default stream) without
 *     synchronising.  *_host functions take HOST pointers, do the copies themselves
 *     and return after the results are in the caller's buffers.
 *   - return value: 0 on success, a negative QMPS_ERR_* code otherwise
 *     (qmps_last_error() gives the message).  Per-problem numerical conditions are
 *     reported in the optional int32 `status` array, never by failing the call:
 *       QMPS_ST_OK          0
 *       QMPS_ST_NOT_PD      1  cholesky(r) failed: the reference raises
 *                              numpy.linalg.LinAlgError here (qmps/tools.py:182)
 *       QMPS_ST_NO_CONVERGE 2  QR iteration hit its sweep limit
 *       QMPS_ST_SINGULAR    3  degenerate leading eigenvalue (fixed point not unique)
 *   - output pointers documented as "optional" may be NULL.
 */
#ifndef QMPS_B200_H
#define QMPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMPS_C128 0
#define QMPS_C64 1

#define QMPS_ST_OK 0
#define QMPS_ST_NOT_PD 1
#define QMPS_ST_NO_CONVERGE 2
#define QMPS_ST_SINGULAR 3

#define QMPS_ERR_ARG (-1)      /* invalid argument */
#define QMPS_ERR_CUDA (-2)     /* CUDA runtime error */
#define QMPS_ERR_UNSUPPORTED (-3)

/* gate codes of an ansatz program (qmps/represent.py:268-423 gate lists as data) */
#define QMPS_G_RZ 0
#define QMPS_G_RX 1
#define QMPS_G_RY 2
#define QMPS_G_H 3
#define QMPS_G_CNOT 4
#define QMPS_G_SWAP 5
#define QMPS_G_CZ 6
#define QMPS_G_XPOW 7
#define QMPS_G_ZZPOW 8
#define QMPS_G_XXPOW 9
#define QMPS_G_YYPOW 10
#define QMPS_G_X 11
#define QMPS_G_Z 12

/* one gate: angle (or exponent) = scale * theta[param] + offset; param < 0: constant */
typedef struct qmps_gate_op {
  int32_t code;
  int32_t q0;     /* target / control / first qubit (qubit 0 = most significant) */
  int32_t q1;     /* CNOT target / second qubit */
  int32_t param;
  double scale;
  double offset;
} qmps_gate_op;

const char* qmps_version(void);
const char* qmps_last_error(void);
/* tuning knobs (no reference counterpart): "d2_pdl" (1: programmatic dependent launch for the
 * D = 2 streaming kernel, default 1), "d2_ctas_per_sm" (resident CTAs per SM of that kernel, default 1 = measured best; 0: occupancy limit),
 * "fp16_fast" (the D = 4 eigenvalue-only kernel: 0 generic shared-memory group kernel, 1-5 register-resident
 * half / quarter-warp forms, 6 / 7 shared-resident half / quarter warp, 8 = 7 with a branch-free reciprocal square
 * root, 9 packed two-kernel form, 10 = 8 with trimmed sweep bodies, default 8 = measured best, see profiles/README.md), "fp64_fast" (the D = 8 eigenvalue-only path: 0 generic
 * CTA-per-problem kernel, 1 one warp per 64 x 64 map, 2 packed two-kernel form, default 2 = measured best),
 * "env_real" (1: the real-form register-resident D = 4, 8 direct solver, default 1; 0: generic kernel). */
int qmps_set_option(const char* name, int value);
/* diagnostics of the D = 4 QR kernel (complex128) on the current device, after a device
 * synchronise: out4 = {problems solved, QR sweeps, problems with a forced deflation, 0}. */
int qmps_debug_counters(unsigned long long* out4, int reset);
/* number of visible CUDA devices (0 if none); never fails */
int qmps_device_count(void);

/* a1  unitary_to_tensor (qmps/tools.py:151-154): U [N][2D][2D] -> A [N][2][D][D] */
int qmps_unitary_to_tensor(int D, int64_t N, const void* U, void* A, int dtype, void* stream);

/* a2  tensor_to_unitary + unitary_extension (qmps/tools.py:123-148, 76-94):
 *     A [N][d][D][D] (left-canonical) -> U [N][dD][dD], U[:, :D] = iso exactly; the
 *     remaining columns are a Householder completion (any orthonormal completion is
 *     valid: the reference's own null_space columns are not unique either). */
int qmps_tensor_to_unitary(int d, int D, int64_t N, const void* A, void* U, int dtype, void* stream);

/* a3  environment_to_unitary (qmps/tools.py:97-108): v [N][n] -> V [N][n][n] with
 *     V[:,0] = v/|v| exactly, Householder completion. */
int qmps_environment_to_unitary(int n, int64_t N, const void* v, void* V, int dtype, void* stream);

/* a4/a5  TransferMatrix(A).eigs() + cholesky (qmps/tools.py:176-182):
 *     in: A [N][d][D][D] (in_is_full_U = 0) or U [N][2D][2D] (in_is_full_U = 1, d = 2).
 *     assume_left_canonical = 1: eta = 1 is known, direct fixed-point solve.
 *     assume_left_canonical = 0: full eigen-solve (any A), r rotated to Hermitian.
 *     out (all optional): eta [N] complex, r [N][D][D] Hermitian trace 1,
 *     C [N][D][D] lower Cholesky factor (r = C C^dagger), status [N]. */
int qmps_env_exact(int d, int D, int64_t N, const void* in, int in_is_full_U,
                   int assume_left_canonical, void* eta, void* r, void* C, int32_t* status,
                   int dtype, void* stream);
/* same with HOST buffers; chunked, copies overlapped with compute */
int qmps_env_exact_host(int d, int D, int64_t N, const void* in, int in_is_full_U,
                        int assume_left_canonical, void* eta, void* r, void* C, int32_t* status,
                        int dtype, int device);
/* a4/a5, D = 2 complex128 left-canonical, PACKED outputs: one 64-byte record per solve
 *     [ r00, Re r01, Im r01, c00, Re c10, Im c10, c11, (double) status ]
 *     (r Hermitian trace 1: r11 = 1 - r00, r10 = conj r01; C lower triangular with real positive diagonal: c01 = 0;
 *     eta = 1 on this path) -- the same information as eta / r / C / status of qmps_env_exact in 64 instead of 148
 *     bytes; the host-buffer path is bound by the PCIe link, so this is what a bandwidth-conscious caller binds.
 *     in: A [N][2][2][2] (in_is_full_U = 0) or U [N][4][4]; packed [N][8] doubles (DEVICE / HOST). */
int qmps_env_exact_packed(int64_t N, const void* in, int in_is_full_U, void* packed, void* stream);
int qmps_env_exact_packed_host(int64_t N, const void* in, int in_is_full_U, void* packed, int device);
/* the scalar drop-in in one call: get_env_exact(U) (qmps/tools.py:176-182) on HOST buffers,
 *     U [N][2D][2D] -> V [N][D^2][D^2] = environment_to_unitary(cholesky(r)), status [N] (optional).
 *     One H2D, three kernels, one D2H, one stream synchronisation. */
int qmps_get_env_exact_host(int D, int64_t N, const void* U, void* V, int32_t* status, int dtype, int device);

/* a6  Map(A,B).right_fixed_point() / .left_fixed_point() (xmps; call sites
 *     qmps/time_evolve_tools.py:87, qmps/loschmidts/time_evo.py:79-82):
 *     A [NA][d][D][D], B [NB][d][D][D].
 *     pair_mode 0: problem i uses A[min(i,NA-1)] , B[min(i,NB-1)], N = max(NA,NB)
 *     pair_mode 1: outer product, problem (ia, ib) -> index ia*NB + ib
 *     pair_mode 2: outer product, B-major output: problem (ia, ib) -> index ib*NA + ia
 *     left = 0: right fixed point (x, r); left = 1: left fixed point (x, l).
 *     out (optional): eta [N] complex, vec [N][D][D] unit Frobenius norm in the gauge of the one
 *     recorded xmps output (Time Evo.ipynb cells 22-24: LAPACK zgeev's convention, the component of
 *     largest modulus real positive),
 *     cost [N] = -sqrt|eta| (a11, qmps/loschmidts/time_evo.py:75-116),
 *     echo [N] = -log|eta|^2, fid [N] = |eta|^2 (a8, qmps/time_evolve_tools.py:84-91),
 *     status [N]. */
int qmps_fixed_point(int d, int D, int64_t NA, const void* A, int64_t NB, const void* B,
                     int pair_mode, int left, void* eta, void* vec, void* cost, void* echo,
                     void* fid, int32_t* status, int dtype, void* stream);
/* same with the phase convention of `vec` selectable: QMPS_GAUGE_ZGEEV (as above) or
 *     QMPS_GAUGE_TRACE: tr(vec) real non-negative (traceless: largest entry real positive) -- a
 *     Hermitian fixed point (A = B) then comes out Hermitian, which the canonical-form routines use. */
#define QMPS_GAUGE_TRACE 0
#define QMPS_GAUGE_ZGEEV 1
int qmps_fixed_point_ex(int d, int D, int64_t NA, const void* A, int64_t NB, const void* B,
                        int pair_mode, int left, int vec_gauge, void* eta, void* vec, void* cost,
                        void* echo, void* fid, int32_t* status, int dtype, void* stream);

/* a7  merge (qmps/time_evolve_tools.py:20-23): A [NA][d1][D][D], B [NB][d2][D][D] ->
 *     M [N][d1*d2][D][D], N = max(NA,NB) (a batch of 1 broadcasts).
 *     W (optional) [NW][d1*d2][d1*d2]: M <- tensordot(W, M, [1,0])
 *     (qmps/loschmidts/time_evo.py:79); with NW > 1 and NA = NB = 1 the output is
 *     one block per gate. */
int qmps_merge(int d1, int d2, int D, int64_t NA, const void* A, int64_t NB, const void* B,
               int64_t NW, const void* W, void* M, int dtype, void* stream);

/* a14 ansatz gate lists (qmps/represent.py:268-423): theta [N][P] (double) -> A
 *     [N][2][D][D] (full_unitary = 0) or U [N][2D][2D] (full_unitary = 1);
 *     nq = log2(D)+1 qubits.  `ops` is a HOST array. */
int qmps_ansatz(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, const double* theta,
                int full_unitary, void* out, int dtype, void* stream);

/* a9 + a12  energy cost (qmps/ground_state.py:150-168, 251-266) with the rotosolve
 *     shift fan-out (qmps/rotosolve.py:175, qmps/tools.py:432-438) fused in:
 *     problem (n, s) evaluates theta[n] + shifts[s] * e_coord.
 *     ops/nops/nq as qmps_ansatz; hmat: DEVICE [4][4] complex; shifts: HOST
 *     [nshift] (NULL / nshift = 0: one unshifted evaluation, coord ignored).
 *     out: energy [N][max(nshift,1)] real, status optional same shape. */
int qmps_energy_theta(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P,
                      const double* theta, const void* hmat, int coord, const double* shifts,
                      int nshift, void* energy, int32_t* status, int dtype, void* stream);
/* energy from tensors: two_site = 0: in = A [N][2][D][D], M = merge(A,A), env of E_AA;
 *     two_site = 1: in = M [N][4][D][D] (merge(A1,A2)), env of E_MM
 *     (qmps/ground_state.py:291-331). */
int qmps_energy_tensor(int D, int64_t N, const void* in, int two_site, const void* hmat,
                       void* energy, int32_t* status, int dtype, void* stream);

/* a12  rotosolve closed forms on DEVICE arrays of costs:
 *     nshift = 3: cost [N][3] at shifts (0, +pi/2, -pi/2) -> theta_star [N]
 *       = -pi/2 - atan2(2 e0 - e+ - e-, e+ - e-) (qmps/rotosolve.py:175) and, if
 *       theta_io is given ([N][P], coordinate `coord`), the wrapped in-place update
 *       of qmps/rotosolve.py:176-177.
 *     nshift = 6: cost [N][6] at shifts (0, pi, pi/2, -pi/2, pi/4, -pi/4) -> fit [N][8]
 *       = (a, b, c, d, P, u, Q, v) of qmps/tools.py:434-447 and theta_star = global
 *       minimiser of P sin(2x+u) + Q sin(x+v) on [-pi, pi].
 *     All arrays double. */
int qmps_rotosolve_fit(int64_t N, int nshift, const double* cost, double* theta_star,
                       double* fit, double* theta_io, int P, int coord, void* stream);

/* a13 exact TFIM Loschmidt rate function loschmidt(t, g0, g1)
 *     (qmps/loschmidts/exact_loschmidt.py:6-20): t [NT] -> out [NT], DEVICE doubles. */
int qmps_loschmidt_rate(int64_t NT, const double* t, double g0, double g1, double* out, void* stream);

/* cfg 5  classical power method (qmps.ipynb cells 29-32): K normalised applications
 *     r <- sum_s A_s r B_s^dagger / |.|_F.  A, B [N][d][D][D]; r_io [N][D][D] (start
 *     vector in, r_K out); rayleigh [N] complex = <r_K, E r_K> (optional). */
/* one UNNORMALISED application of the (mixed) transfer map, Y [N][D][D] = sum_s A_s X B_s^dagger (X != Y): the body of
 *     the power-method loop of qmps.ipynb cells 29-32 without its normalisation, and the building block of the Neumann /
 *     Krylov iterations a large-D caller runs (iMPS.dA_dt at D > 16: scripts/classical_time_evolution.py:22-26). */
int qmps_tm_apply(int d, int D, int64_t N, const void* A, const void* B, const void* X, void* Y, int dtype, void* stream);
int qmps_tm_power(int d, int D, int64_t N, const void* A, const void* B, void* r_io, int K,
                  void* rayleigh, int dtype, void* stream);

/* cfg 5  building block of the complex64 power method: batched complex product on the tcgen05
 *     tensor cores (kind::tf32, 3xTF32 split = FP32-grade results),
 *         C[b] = sum_t X[b][t] . op(Y[b][t]),   op = transpose (conj_y = 0) / conjugate transpose (1)
 *     X [batch][nsum][M][K], Y [batch][nsum][N][K] (both K-major), C [batch][M][N], complex64,
 *     DEVICE pointers; M and N multiples of 64, K a multiple of 32 (else QMPS_ERR_UNSUPPORTED).
 *     Stage 1 (A_s . r) and stage 2 (sum_s T_s . B_s^dagger) of qmps_tm_power are two calls of the
 *     same kernel (qmps.ipynb cells 29-32). */
int qmps_cgemm_c64_tc(int64_t batch, int nsum, int M, int N, int K, const void* X, const void* Y,
                      int conj_y, void* C, void* stream);

/* the complex128 counterpart on tcgen05 kind::i8: C[b] = X[b] . Y[b]^T (conj_y = 0) or X[b] . Y[b]^H (1) with
 *     X [batch][M][K], Y [batch][N][K], C [batch][M][N] complex128, M % 64 == 0, N % 32 == 0, K % 64 == 0.
 *     Every real operand row is cut into six signed 7-bit slices under a power-of-two row scale; the 21 slice
 *     products with i + j < 6 are exact int32 tensor-core products, recombined in FP64 (relative error ~4e-12).
 *     qmps_tm_power uses it for complex128 when D % 64 == 0 (option "i8_power": 0 = FP64 tensor pipe instead). */
int qmps_zgemm_c128_i8(int64_t batch, int M, int N, int K, const void* X, const void* Y, int conj_y, void* C,
                       void* stream);


/* ---- canonical forms and local expectation values (xmps' iMPS methods as the
 *      reference's loops call them; xmps is not vendored, so the gauge is this build's documented
 *      Cholesky gauge -- every quantity the call sites consume is gauge invariant) ---------------- */

/* iMPS([A]).left_canonicalise() (call sites qmps/time_evolve_tools.py:85-86,
 *     qmps/loschmidts/time_evo.py:76,143; scripts/loschmidt.py:210,368):
 *     A [N][d][D][D] (any normalisable tensor) -> AL = L A L^-1 / sqrt(eta) with
 *     sum_s AL_s^dagger AL_s = 1, where l = L^dagger L is the left fixed point of E_AA
 *     (L upper triangular, positive diagonal, tr(l) = D).
 *     out: AL [N][d][D][D]; optional eta [N] complex, L [N][D][D], status [N]
 *     (QMPS_ST_NOT_PD when l is not positive definite, QMPS_ST_NO_CONVERGE from the eigen-solve). */
int qmps_left_canonicalise(int d, int D, int64_t N, const void* A, void* AL, void* eta, void* L,
                           int32_t* status, int dtype, void* stream);

/* iMPS([A]).mixed() -> (AL, AR, C) (qmps/tools.py:184-186 get_env_exact_alternative,
 *     qmps/ground_state.py:287, tests/test_represent.py:18-31):
 *     AL as above (assume_left_canonical = 1: AL = A, no eigen-solve), C lower Cholesky factor of the
 *     trace-1 right fixed point r of E_ALAL (r = C C^dagger), AR = C^-1 AL C, so that
 *     Map(AL,AL): right r, left 1;  Map(AR,AR): right 1, left C^dagger C   (test_represent.py:23-31).
 *     All outputs optional: AL, AR [N][d][D][D], C [N][D][D], eta [N], status [N]. */
int qmps_mixed_canonical(int d, int D, int64_t N, const void* A, int assume_left_canonical, void* AL,
                         void* AR, void* C, void* eta, int32_t* status, int dtype, void* stream);

/* building block of the two above: A' = L A L^-1 / sqrt|eta| with X = l Hermitian PD, l = L^dagger L
 *     (x_kind 0; G_out = L), or A' = C^-1 A C with X = C lower triangular (x_kind 1).
 *     eta, G_out, status optional. */
int qmps_gauge_transform(int d, int D, int64_t N, const void* A, const void* X, int x_kind,
                         const void* eta, void* A_out, void* G_out, int32_t* status, int dtype,
                         void* stream);

/* iMPS([A]).Es(ops) / .E(op) (qmps/loschmidts/time_evo.py:144, scripts/loschmidt.py:369,
 *     tests/test_represent.py:37): single-site expectation values
 *       out[n][o] = sum_st ops[o][s][t] tr(l A_t r A_s^dagger) / (eta tr(l r)).
 *     A [N][d][D][D], r [N][D][D] right fixed point; lvec (optional) = `vec` of
 *     qmps_fixed_point(left = 1) and eta (optional) for a tensor that is not left-canonical;
 *     lvec = eta = NULL: A left-canonical, tr r = 1.  ops DEVICE [nops][d][d]; out [N][nops] complex. */
int qmps_expectation(int d, int D, int64_t N, const void* A, const void* r, const void* lvec,
                     const void* eta, int nops, const void* ops, void* out, int dtype, void* stream);

/* ---- two-layer brick-wall iMPS family (new_tdvp/ClassicalTDVPStripped.py) --------------
 * Unitaries are 4x4, row-major: U[(a,b),(c,d)] = U.reshape(2,2,2,2)[a,b,c,d].  (U1, U2) is the ket
 * state, (U1_, U2_) what the reference passes as `U1_`, `U2_` -- ALREADY daggered by the caller -- unless
 * bra_undaggered = 1: then the arrays hold the candidate unitaries V1, V2 themselves (what paramU returns,
 * :159-180) and the daggers are formed on the device.  Every array count (NK, NB, NM, NO, NW) is N or 1
 * (1 = shared by the whole batch). */

/* RightEnvironment / LeftEnvironment .exact_environment_circuit + .exact_environment (:322-352, :394-426):
 *     side 0 = right, 1 = left.  Optional out: mat [N][4][4] the 4x4 map; eta [N] its eigenvalue selected
 *     by numpy's complex argmax (lexicographic on (real, imag) -- the reference's rule, not the largest
 *     modulus); vec [N][2][2] the eigenvector in scipy.linalg.eig's (zgeev) gauge: unit 2-norm, component
 *     of largest modulus real positive; status [N] (QMPS_ST_NO_CONVERGE). */
int qmps_bw_environment(int side, int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB,
                        const void* U1_, const void* U2_, int bra_undaggered, void* mat, void* eta,
                        void* vec, int32_t* status, int dtype, void* stream);

/* RightEnvironment.circuit(U1, U2, U1_, U2_, M) (:360-384): one application of the right map to
 *     M [NM][2][2] -> out [N][2][2]. */
int qmps_bw_env_apply(int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB, const void* U1_,
                      const void* U2_, int bra_undaggered, int64_t NM, const void* M, void* out, int dtype,
                      void* stream);

/* OverlapCalculator.expectation_value(U1, U2, O) (:428-533): Re <psi| 1 (x) O (x) 1 |psi> with
 *     op_qubits = 2 (O [NO][4][4], 4-qubit state) or 4 (O [NO][16][16], 6-qubit state); out [N] real. */
int qmps_bw_expectation(int64_t N, int64_t NK, const void* U1, const void* U2, int op_qubits, int64_t NO,
                        const void* O, void* out, int dtype, void* stream);

/* ManifoldOverlap.circuit(U1, U2, U1_, U2_, Mr, Ml, W) (:228-268): <phi| Ml (x) W (x) Mr |psi> on six
 *     qubits; Mr, Ml [NM][2][2], W [NW][16][16]; overlap [N] complex. */
int qmps_bw_overlap(int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB, const void* U1_,
                    const void* U2_, int bra_undaggered, int64_t NM, const void* Mr, const void* Ml, int64_t NW,
                    const void* W, void* overlap, int dtype, void* stream);

/* body of Evolve.exact_cost_function (:777-790) for candidate unitaries V1, V2 [NB][4][4] (undaggered):
 *     Mr = exact right environment of the mixed map, overlap with (Mr, Mr^dagger) and W, cost = -|overlap|^2,
 *     all in one launch.  cost [N] real; optional overlap [N], eta [N], Mr [N][2][2], status [N]. */
int qmps_bw_evolve_cost(int64_t N, int64_t NK, const void* U1, const void* U2, int64_t NB, const void* V1,
                        const void* V2, int64_t NW, const void* W, void* cost, void* overlap, void* eta, void* Mr,
                        int32_t* status, int dtype, void* stream);

/* (e)  local part of the final cost reduction: (min cost, argmin + index_offset) of a
 *     DEVICE array, written to DEVICE best_cost[1] / best_index[1]; the cross-rank
 *     step is one NCCL all-gather of 16 bytes per rank (qmps_b200/dist.py). */
int qmps_argmin(int64_t N, const double* cost, int64_t index_offset, double* best_cost,
                int64_t* best_index, void* stream);


/* ---- single-call pipelines (SURVEY 8(b) proposal) ---------------------------------------------------- */

/* a11 in one call (qmps/loschmidts/time_evo.py:75-116 = scripts/loschmidt.py:209-239 for a whole grid):
 *     for NP parameter vectors theta [NP][P] of the gate program and NT two-site gates W [NT][4][4] applied to
 *     the state tensor A0 [2][D][D] (D = 2^(nq-1)):  eta_2 = leading eigenvalue of
 *     Map(W_k . merge(A0,A0), merge(B_p,B_p)), cost [NP][NT] = -sqrt|eta_2|, echo [NP][NT] = -log|eta_2|^2,
 *     eta [NP][NT] complex, status [NP][NT]; any output may be NULL.  Four launches on `stream`. */
int qmps_loschmidt_batched(const qmps_gate_op* ops, int nops, int nq, int64_t NP, int P, const double* theta,
                           const void* A0, int64_t NT, const void* W, void* cost, void* echo, void* eta,
                           int32_t* status, int dtype, void* stream);
/* same with HOST buffers (pageable or pinned): 8 P bytes in and 8 NT bytes out per parameter vector */
int qmps_loschmidt_batched_host(const qmps_gate_op* ops, int nops, int nq, int64_t NP, int P, const double* theta,
                                const void* A0, int64_t NT, const void* W, void* cost, void* echo, void* eta,
                                int32_t* status, int dtype, int device);

/* a9 / a12 with HOST buffers: theta [N][P] in, energy [N][max(nshift,1)] (+ status) out
 *     (qmps/ground_state.py:150-168, 251-266; shifts as qmps/rotosolve.py:175) */
int qmps_energy_theta_host(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, const double* theta,
                           const void* hmat, int coord, const double* shifts, int nshift, void* energy,
                           int32_t* status, int dtype, int device);

/* a12 whole coordinate sweeps on the device: for each of n_sweeps sweeps and each coordinate i the fused
 *     shift fan-out + energy launch and the closed-form update (two_frequency = 0: qmps/rotosolve.py:154-181,
 *     3 shifts; 1: qmps/tools.py:422-457, 6 shifts).  theta [N][P] (DEVICE) is updated in place; energy [N]
 *     (optional, DEVICE) receives the cost after the last sweep.  No host round trip. */
int qmps_rotosolve_sweep(const qmps_gate_op* ops, int nops, int nq, int64_t N, int P, double* theta, const void* hmat,
                         int n_sweeps, int two_frequency, void* energy, int dtype, void* stream);

/* (f)-2 the reference's time-evolution loop on the device (scripts/loschmidt.py:367-375 = qmps/loschmidts/time_evo.py:140-150):
 *     theta_traj[0] = theta0; for each of n_steps steps  theta_traj[t+1] = argmin_p obj(p, A(theta_traj[t]), W)  with
 *     obj(p, A, W) = -sqrt|eta(Map(W . merge(A,A), merge(B_p,B_p)))| (qmps/loschmidts/time_evo.py:75-116).  The reference
 *     calls scipy.optimize.minimize per step; here each step is n_gen generations of a population search -- npop candidate
 *     vectors per generation evaluated by ONE qmps_loschmidt_batched call, the argmin fed back on the device, step size
 *     adapted from sigma0 -- followed by n_bfgs iterations of BFGS (what scipy's default minimiser does) whose
 *     finite-difference gradient (2 P + 1 points) and line search (24 step sizes) are one launch each; no host round
 *     trip anywhere.  n_gen = 0 starts BFGS from theta_t itself.  P <= 64.  Outputs (DEVICE): theta_traj [(n_steps+1)][P] doubles,
 *     step_cost [n_steps] (the minimum found per step; 1 + step_cost is the projection error), echo [n_steps+1] =
 *     |eta(E_{A_t A_0})|^2 (``A_.overlap(A)``; the reference plots -log of it against exact_loschmidt). */
int qmps_loschmidt_trajectory(const qmps_gate_op* ops, int nops, int nq, int P, const double* theta0, const void* W,
                              int n_steps, int n_gen, int npop, double sigma0, uint64_t seed, int n_bfgs, double* theta_traj,
                              void* step_cost, void* echo, int dtype, void* stream);

/* (f)-4 PXP scar dynamics (the reference's scars.py, a caller of Map(merge(A1,A2), merge(A1',A2')).right_fixed_point()):
 *     cost [N] = -2 |<0|C|0>| of scars_time_evolve_cost_function / scars_cost_fun_alternate (scars.py:76-155) for
 *     candidate angles params [N][4] = [theta1, phi1, phi2, theta2] against current [NC][4] (NC = 1: shared, or N);
 *     W: DEVICE [16][16] complex, the four-site gate expm(+i dt H(mu)) (scars.py:23-29); eta [N] (optional): the
 *     leading eigenvalue of the mixed two-site map.  trajectory: simulate_scars (scars.py:157-170) on the device --
 *     traj [n_steps+1][4] DEVICE (row 0 = params0), step_cost [n_steps] optional. */
int qmps_scars_cost(int64_t N, const double* params, int64_t NC, const double* current, const void* W, void* cost,
                    void* eta, int32_t* status, int dtype, void* stream);
int qmps_scars_trajectory(const double* params0, const void* W, int n_steps, int n_gen, int npop, double sigma0,
                          uint64_t seed, int n_bfgs, double* traj, void* step_cost, int dtype, void* stream);

/* (e)  the final cost reduction across ranks: local argmin, one ncclAllGather of 16 bytes per
 *     rank and a final pass, all on `stream`; best_cost [1] / best_index [1] are DEVICE scalars, identical on
 *     every rank (ties: smallest global index; an empty shard contributes nothing).  `comm` is an ncclComm_t --
 *     the caller's own (qmps_nccl_comm_create) or an existing one, e.g. torch's ProcessGroupNCCL communicator;
 *     NULL means a single rank.  NCCL is bound at run time to the libnccl.so.2 already loaded in the process. */
int qmps_argmin_allreduce(void* comm, int64_t N, const double* cost, int64_t index_offset, double* best_cost,
                          int64_t* best_index, void* stream);
int qmps_nccl_unique_id(void* id128);                                        /* ncclGetUniqueId (128 bytes), rank 0 */
int qmps_nccl_comm_create(const void* id128, int world, int rank, void** comm);   /* ncclCommInitRank on the current device */
int qmps_nccl_comm_destroy(void* comm);


/* (f)-3 classical iTDVP of single-site uniform MPS, batched over N independent states (xmps iMPS.dA_dt /
 *     iTDVP.Trajectory; call sites scripts/classical_time_evolution.py:16-27, qmps/loschmidts/mps_loschmidts.py:20-22,
 *     scripts/mixed_environment.py:41).  h [d*d][d*d] is the two-site Hamiltonian (DEVICE), D <= 16 (any), d <= 4.
 *     imaginary = 1 replaces -i by -1 (imaginary-time flow towards the ground state).
 *     tangent: AL [N][d][D][D] LEFT-CANONICAL -> dA [N][d][D][D], energy [N] (real, <h> per bond), status [N].
 *     dadt:    A in ANY gauge (canonicalise, tangent, transform back: the left gauge condition is covariant). */
int qmps_tdvp_tangent(int d, int D, int64_t N, const void* AL, const void* h, int imaginary, void* dA, void* energy,
                      int32_t* status, int dtype, void* stream);
int qmps_tdvp_dadt(int d, int D, int64_t N, const void* A, const void* h, int imaginary, void* dA, void* energy,
                   int32_t* status, int dtype, void* stream);
/* n_steps steps of size dt from A_io (replaced by its canonical form, then by the final state), no host round trip:
 *     method 1 = the reference's RK4 loop (classical_time_evolution.py:22-26, left_canonicalise after each step),
 *     method 0 = explicit Euler (Trajectory.eulerint).  Optional outputs: traj [n_steps+1][N][d][D][D],
 *     rates [n_steps+1][N] = -log|eta(E_{A_t A_0})|^2 (Trajectory.loschmidts()), energy [n_steps][N] (at the start
 *     of each step). */
int qmps_tdvp_evolve(int d, int D, int64_t N, void* A_io, const void* h, double dt, int n_steps, int method, int imaginary,
                     void* traj, void* rates, void* energy, int32_t* status, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QMPS_B200_H */
