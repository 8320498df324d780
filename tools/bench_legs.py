"""The legs of the metric beyond the headline: BASELINE.json's metric is "environment solves/sec AND
Loschmidt-echo steps/sec at D = 2 / 8 / 64"; bench.py's headline is config 2 (D = 2 solves) and every other cell
is a `sub_results` entry built here.  Each leg measures, on the same synthetic inputs SURVEY 8(d) fixes:

  value         device-timed throughput through the C ABI, inputs resident in HBM (CUDA events on the launch stream)
  e2e           the same through the host-buffer entry / the numpy-facing Python call, copies inside the timed region
  roofline      algorithmic flops (or bytes) per unit x units per step / kernel time, against the MEASURED peak of the
                pipe the kernel issues to (profiles/peaks_r0*.json from tools/peaks.cu; MEASURED_PEAKS.json for HBM)
  cpu_baseline  the per-call oracle port on all host cores ("port"), and the B2 variant of BASELINE.md section 3 --
                the same algorithm as stacked / threaded numpy ("vectorised") -- each on a bounded sample

Algorithmic work per unit: dense complex eigenvalues of an n x n map (Hessenberg + shifted QR,
what numpy.linalg.eig does) = 100 n^3 real flops (n = 4: 6.4e3, n = 16: 4.1e5); direct environment solve
8 d D^4 + (8/3) D^6; one transfer-matrix application 32 D^3 (d = 2).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_peaks():
    """Measured pipe peaks (tools/peaks.cu on a B200 of this pool) + the driver's HBM / bf16 figures."""
    peaks = {"fp64_fma_tflops": 33.86, "fp64_dmma_tflops": 37.03, "fp32_fma_tflops": 69.45, "source": "fallback (profiles/peaks_r01.json values)"}
    for name in ("peaks_r02.json", "peaks_r01.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                peaks.update(json.load(f))
            peaks["source"] = f"profiles/{name} (tools/peaks.cu, measured)"
            break
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            m = json.load(f)
        peaks["hbm_gbs"] = m.get("hbm_gbs", 6515.4)
        peaks["bf16_tflops"] = m.get("bf16_tflops")
        peaks["hbm_source"] = "MEASURED_PEAKS.json"
    else:
        peaks.setdefault("hbm_gbs", 6650.0)
        peaks["hbm_source"] = "fallback 6.65 TB/s (B200_PROFILING.md)"
    return peaks


def timed_ms(torch, fn, reps=3, warm=1):
    """median ms of fn() over `reps` timed calls (CUDA events on the current stream, sync on both sides)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def wall_ms(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def tfim_gates(NT, dt=0.02, g=0.2):
    from scipy.linalg import expm
    from qmps_b200.ground_state import Hamiltonian
    H = Hamiltonian({'ZZ': -1, 'X': g}).to_matrix()
    return np.stack([expm(-1j * H * 2 * dt * k) for k in range(NT)])


# ---------------------------------------------------------------------------------------------------------------
# CPU baselines: bounded samples on all host cores
# ---------------------------------------------------------------------------------------------------------------
def _one_thread():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:  # noqa: BLE001
        pass


def _port_worker(args):
    """per-call oracle evaluations for `seconds` on one core -> (count, seconds)"""
    kind, seconds, seed = args
    _one_thread()
    import oracle as O
    from scipy.stats import unitary_group
    rng = np.random.default_rng(seed)
    if kind == "env_d2":
        Us = [unitary_group.rvs(4, random_state=seed * 131 + k) for k in range(64)]
        fn = lambda k: O.env_exact_parts(O.unitary_to_tensor(Us[k & 63]))          # (eta, r, C, V[:,0])
    elif kind in ("loschmidt_d2", "loschmidt_d4", "loschmidt_d8"):
        D = int(kind[-1])
        tens = (lambda p: O.unitary_to_tensor(O.shallow_full_state_tensor(p))) if D == 2 else \
               (lambda p: O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(D, p)))
        P = {2: 15, 4: 12, 8: 24}[D]
        A0 = tens(rng.normal(size=P))
        W = tfim_gates(2)[1]
        thetas = rng.normal(size=(16, P))
        fn = lambda k: O.loschmidt_cost(A0, tens(thetas[k & 15]), W)                # gate(theta) -> tensor -> cost
    elif kind == "energy_d8":
        H = O.heisenberg_matrix()
        thetas = rng.normal(size=(16, 24))
        fn = lambda k: O.energy_transfer(O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(8, thetas[k & 15])), H)
    elif kind in ("power_d64", "power_d256"):
        D = 64 if kind.endswith("64") else 256
        A = O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=4))
        Bt = O.unitary_to_tensor(unitary_group.rvs(2 * D, random_state=5))
        state = {"r": np.eye(D, dtype=complex) / np.sqrt(D)}

        def fn(k):
            r = sum(A[s] @ state["r"] @ Bt[s].conj().T for s in range(2))
            state["r"] = r / np.linalg.norm(r)
    else:
        raise ValueError(kind)
    fn(0)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn(n); n += 1
    return n, time.perf_counter() - t0


def _vec_worker(args):
    """stacked-numpy evaluations for `seconds` on one core -> (units, seconds)"""
    kind, seconds, seed = args
    _one_thread()
    import oracle as O
    rng = np.random.default_rng(seed)
    if kind == "env_d2":
        A = O.tensors_of_unitaries(O.haar_unitaries(4, 8192, seed))

        def fn():
            eta, r = O.stacked_env_exact(A)
            O.stacked_cholesky_env(r)
            return len(A)
    elif kind in ("loschmidt_d2", "loschmidt_d4"):
        D = 2 if kind.endswith("d2") else 4
        Bs = O.tensors_of_unitaries(O.haar_unitaries(2 * D, 64, seed))
        A0 = Bs[0]
        Ws = tfim_gates(16)

        def fn():
            O.stacked_loschmidt_costs(A0, Bs, Ws)
            return 64 * 16
    elif kind == "energy_d8":
        A = O.tensors_of_unitaries(O.haar_unitaries(16, 64, seed))     # tensors given: theta -> U is not vectorised
        H = O.heisenberg_matrix()

        def fn():
            O.stacked_energy_transfer(A, H)
            return 64
    else:
        raise ValueError(kind)
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        n += fn()
    return n, time.perf_counter() - t0


def cpu_rate(pool, cores, kind, seconds, vectorised=False):
    res = pool.map(_vec_worker if vectorised else _port_worker, [(kind, seconds, 17 + c) for c in range(cores)])
    return sum(n / dt for n, dt in res)


def cpu_power_threaded(D, N, seconds):
    """B2 for config 5: the application r <- sum_s A_s r B_s^dagger as batched matmuls on the threaded BLAS of the
    box (all cores of one process)."""
    import oracle as O
    A = np.stack([O.unitary_to_tensor(u) for u in O.haar_unitaries(2 * D, min(N, 4), 4)])
    A = np.tile(A, (max(1, N // len(A)), 1, 1, 1))[:N]
    Bh = np.ascontiguousarray(A.conj().transpose(0, 1, 3, 2))
    r = np.tile(np.eye(D, dtype=complex) / np.sqrt(D), (N, 1, 1))

    def fn():
        nonlocal r
        x = sum(np.matmul(np.matmul(A[:, s], r), Bh[:, s]) for s in range(2))
        r = x / np.linalg.norm(x, axis=(1, 2), keepdims=True)
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        fn(); n += N
    return n / (time.perf_counter() - t0)


def cpu_tdvp_threaded(D, seconds, tol=1e-11):
    """CPU arm of the large-D TDVP tangent leg: the SAME iterative algorithm as batched.tdvp_tangent_large (power method
    for r, Neumann series of the left transfer map for K, the formulas of oracle/tdvp.py:59-83) in numpy on the threaded
    BLAS of the box (all cores, one tensor at a time) -- the dense D^2 x D^2 solve of the oracle takes ~10 s per tangent at
    D = 64 and is not what a CPU user would run.  Returns tangents/s."""
    rng = np.random.default_rng(64)
    d = 2
    Q = np.linalg.qr(rng.normal(size=(d * D, D)) + 1j * rng.normal(size=(d * D, D)))[0]
    AL = np.ascontiguousarray(Q.reshape(D, d, D).transpose(1, 0, 2))
    h = (-np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0])) + 0.35 * (np.kron([[0, 1], [1, 0]], np.eye(2)) + np.kron(np.eye(2), [[0, 1], [1, 0]]))).reshape(d, d, d, d).astype(complex)
    AH = np.ascontiguousarray(AL.conj().transpose(0, 2, 1))

    def one():
        r, prev = np.eye(D, dtype=complex) / D, None
        for it in range(4096):
            x = sum(AL[s] @ r @ AH[s] for s in range(d))
            ray = np.vdot(r, x).real / np.vdot(r, r).real
            r = x / np.linalg.norm(x)
            if it % 32 == 31:
                if prev is not None and abs(ray - prev) <= tol:
                    break
                prev = ray
        r = 0.5 * (r + r.conj().T); r = r / np.trace(r).real
        AA = np.einsum("sij,tjk->stik", AL, AL)
        C = np.einsum("abcd,cdik->abik", h, AA)
        Hl = np.einsum("stji,stjk->ik", AA.conj(), C)
        e = np.einsum("ik,ki->", Hl, r).real
        Bm = Hl - e * np.eye(D)
        K = Bm.copy()
        for _ in range(4096):
            Kn = Bm + sum(AH[s] @ K @ AL[s] for s in range(d))
            done = np.abs(Kn - K).max() <= tol * np.abs(Kn).max()
            K = Kn
            if done:
                break
        K = K - np.einsum("ik,ki->", K, r) * np.eye(D)
        rinv = np.linalg.inv(r)
        G = (np.einsum("stik,kl,tml,mj->sij", C, r, AL.conj(), rinv, optimize=True) + np.einsum("tki,tskj->sij", AL.conj(), C, optimize=True)
             + np.einsum("ik,skj->sij", K, AL))
        P = np.einsum("ski,skj->ij", AL.conj(), G)
        return -1j * (G - np.einsum("sik,kj->sij", AL, P))
    one()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        one(); n += 1
    return n / (time.perf_counter() - t0)


def cpu_baselines(kinds, seconds=2.0, cores=None):
    """{kind: {"port": rate, "vectorised": rate or None}} on `cores` host cores."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    out = {}
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_port_worker, [("env_d2", 0.2, c) for c in range(cores)])      # imports, page-in
        for kind in kinds:
            if kind == "tdvp_d64":
                continue
            port = cpu_rate(pool, cores, kind, seconds)
            vec = None
            if kind in ("env_d2", "loschmidt_d2", "loschmidt_d4", "energy_d8"):
                vec = cpu_rate(pool, cores, kind, seconds, vectorised=True)
            out[kind] = {"port": port, "vectorised": vec}
    for kind in kinds:
        if kind in ("power_d64", "power_d256"):
            D = 64 if kind.endswith("64") else 256
            out[kind]["vectorised"] = cpu_power_threaded(D, 64 if D == 64 else 8, seconds)
    if "tdvp_d64" in kinds:
        out["tdvp_d64"] = {"port": cpu_tdvp_threaded(64, seconds), "vectorised": None}
    return out, cores


# ---------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------
def _roof(bound, achieved, peak, unit, kernel, per_unit, note=None, traffic=None):
    d = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak if peak else None,
         "kernel": kernel, "algorithmic_per_unit": per_unit, "traffic": traffic}
    if note:
        d["note"] = note
    return d


def leg_loschmidt(torch, B, R, dev, D, peaks, scale=1.0, c64_too=True):
    """cfg 7 (D = 2) / cfg 3 (D = 4): 4096 parameter sets x 1000 times, one qmps_loschmidt_batched call; D = 8 (the metric's
    middle bond dimension; 64 x 64 mixed maps on the packed warp-per-problem kernels): 256 parameter sets x 100 times."""
    NP, NT = max(64, int(4096 * scale)), 1000
    if D == 2:
        P, seed, gate = 15, 7, R.ShallowFullStateTensor(2, np.zeros(15))
    elif D == 8:
        NP, NT = max(32, int(256 * scale)), 100
        P, seed, gate = 24, 8, R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24))
    else:
        P, seed, gate = 12, 2, R.ShallowCNOTStateTensor_nonuniform(4, np.zeros(12))
    prog = gate.program()
    theta_h = np.random.default_rng(seed).normal(size=(NP, P))
    theta = torch.from_numpy(theta_h).to(dev)
    A0 = B.ansatz_tensors(prog, theta[:1])[0]
    A0_h = A0.cpu().numpy()
    W_h = tfim_gates(1000)[:NT].copy()
    W = torch.from_numpy(W_h).to(dev)
    ms = timed_ms(torch, lambda: B.loschmidt_costs(prog, theta, A0, W), reps=3, warm=1)
    units = NP * NT
    n = D * D
    flops = 100.0 * n ** 3
    res = {"cfg": {2: 7, 4: 3, 8: 38}[D], "workload": f"loschmidt_D{D}_{NP}x{NT}_c128",
           "metric": "loschmidt_echo_steps_per_sec", "unit": "steps/s", "dtype": "c128",
           "value": units / ms * 1e3, "ms_per_step": ms, "units_per_step": units,
           "api": "qmps_loschmidt_batched (ansatz, merge, gate-merge, fixed points; 4 launches)",
           "roofline": _roof("fp64", units * flops / ms * 1e3 / 1e12, peaks["fp64_fma_tflops"], "TFLOP/s",
                             {2: "fp_d2_kernel<double>", 4: "fp16s8_kernel<double>", 8: "fp64p_hess_kernel<double> + fp64p_qr_kernel<double>"}[D], f"{flops:.3g} real flops (100 n^3, n = {n})",
                             note="eigenvalues of every map by Hessenberg + shifted QR; iteration counts are data dependent, "
                                  "the algorithmic count is the LAPACK-style estimate")}
    if c64_too:
        ms32 = timed_ms(torch, lambda: B.loschmidt_costs(prog, theta, A0, W, dtype=torch.complex64), reps=3, warm=1)
        res["value_c64"] = units / ms32 * 1e3
    # e2e: host arrays in, host arrays out (8 P bytes in, 16 NT bytes out per parameter set)
    oc = torch.empty((NP, NT), dtype=torch.float64).pin_memory().numpy()
    oe = torch.empty((NP, NT), dtype=torch.float64).pin_memory().numpy()
    ms_e = wall_ms(lambda: B.loschmidt_costs_host(prog, theta_h, A0_h, W_h, out_cost=oc, out_echo=oe), reps=2, warm=1)
    ms_pg = wall_ms(lambda: B.loschmidt_costs_host(prog, theta_h, A0_h, W_h), reps=2, warm=1)
    res["e2e"] = {"value": units / ms_e * 1e3, "unit": "steps/s", "h2d_bytes_per_step": theta_h.nbytes + W_h.nbytes + A0_h.nbytes,
                  "d2h_bytes_per_step": 2 * 8 * units, "api": "qmps_loschmidt_batched_host (inputs pageable numpy, cost / echo into pinned host arrays)",
                  "pageable_outputs_value": units / ms_pg * 1e3}
    return res


def leg_energy_d8(torch, B, R, dist_mod, dev, peaks, rank=0, world=1, scale=1.0):
    """cfg 4: 2^16 parameter vectors of the D = 8 Heisenberg ansatz, 3 rotosolve shifts of one coordinate each
    (196 608 energy evaluations), the closed-form update, and the global (min E, argmin) -- STRONG scaling: the
    2^16 vectors are split over the ranks and the 16-byte-per-rank NCCL all-gather is inside the timed region."""
    from qmps_b200.ground_state import Hamiltonian
    from qmps_b200 import dist as D_
    N = max(256, int(65536 * scale))
    lo, hi = D_.shard_range(N, rank, world)
    theta_all = np.random.default_rng(3).normal(size=(N, 24))
    theta_h = np.ascontiguousarray(theta_all[lo:hi])
    theta = torch.from_numpy(theta_h).to(dev)
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
    H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
    out = {}

    def step():
        e = B.energy_theta(prog, theta, H, coord=5, shifts=B.ROTO3_SHIFTS)
        out["e"] = e
        out["best"] = D_.argmin_allreduce(e[:, 0].contiguous(), index_offset=lo)

    def step_no_coll():
        out["e"] = B.energy_theta(prog, theta, H, coord=5, shifts=B.ROTO3_SHIFTS)
    ms = timed_ms(torch, step, reps=3, warm=1)
    ms_nc = timed_ms(torch, step_no_coll, reps=3, warm=1)
    return {"ms": ms, "ms_without_collective": ms_nc, "N": N, "n_local": hi - lo, "evals_local": 3 * (hi - lo),
            "theta_h": theta_h, "prog": prog, "H": H, "best_cost": float(out["best"][0].item()), "best_index": int(out["best"][1].item())}


def finish_energy_d8(torch, B, raw, ms_max, world, peaks):
    N = raw["N"]
    units = 3 * N
    flops = 8 * 2 * 8 ** 4 + (8.0 / 3.0) * 8 ** 6
    res = {"cfg": 4, "workload": f"rotosolve3_energy_D8_{N}_vectors_c128", "metric": "energy_evaluations_per_sec", "unit": "evals/s",
           "dtype": "c128", "value": units / ms_max * 1e3, "ms_per_step": ms_max, "units_per_step": units, "n_gpus": world,
           "scaling": "strong", "collective": "qmps_argmin_allreduce (local argmin + 16 B/rank ncclAllGather + final pass) inside the timed region",
           "collective_share": max(0.0, 1.0 - raw["ms_without_collective"] / raw["ms"]),
           "best": {"cost": raw["best_cost"], "index": raw["best_index"]},
           "api": "qmps_energy_theta (ansatz kernel with the shift fan-out -> A through HBM -> 64 x 64 real solve on the FP64 tensor pipe -> energy) + qmps_argmin_allreduce",
           "roofline": _roof("fp64", (units / world) * flops / raw["ms_without_collective"] * 1e3 / 1e12, peaks["fp64_fma_tflops"], "TFLOP/s",
                             "ansatz_kernel<double,32> + env_dmma_kernel<1,2>", f"{flops:.4g} real flops (8 d D^4 + (8/3) D^6: the complex LU the reference's dense route implies; the kernel solves the equivalent 64 x 64 REAL system, ~1/3 of those flops)")}
    return res


def leg_power(torch, B, dev, D, peaks, cdt_name, scale=1.0, nprob=None, with_e2e=True):
    """cfg 5: K = 32 normalised applications r <- sum_s A_s r B_s^dagger on N problems (512 at D = 64, 32 at 256)."""
    cdt = torch.complex128 if cdt_name == "c128" else torch.complex64
    N = nprob or max(2, int((512 if D == 64 else 32) * scale))
    K = 32
    g = torch.Generator(device=dev).manual_seed(4)

    def lc(seed):
        g.manual_seed(seed)
        Z = torch.randn((N, 2 * D, D), dtype=torch.float64, device=dev, generator=g) + \
            1j * torch.randn((N, 2 * D, D), dtype=torch.float64, device=dev, generator=g)
        Q, _ = torch.linalg.qr(Z)
        return Q.reshape(N, D, 2, D).permute(0, 2, 1, 3).contiguous().to(cdt)
    A, Bt = lc(4), lc(5)
    ms = timed_ms(torch, lambda: B.tm_power(A, Bt, K), reps=3, warm=1)
    apps = N * (K + 1)
    flops = 32.0 * D ** 3
    achieved = apps * flops / ms * 1e3 / 1e12
    i8 = cdt_name == "c128" and D >= 64 and D % 64 == 0            # qmps_tm_power's default dispatch (option i8_power = 1)
    if i8:
        peak, pipe, kern = peaks.get("i8_tcgen05_tops") or 2.0 * (peaks.get("bf16_tflops") or 1620.5), "tcgen05 kind::i8 (TOP/s)", "zgemm_i8_kernel"
        issued, what = achieved * 21.0, "; 21 exact int8 slice products issued per algorithmic real product"
    elif cdt_name == "c128":
        peak, pipe, kern = peaks["fp64_dmma_tflops"], "fp64 tensor (DMMA)", "zgemm_dmma_kernel"
        issued, what = achieved, ""
    else:
        peak, pipe, kern = peaks.get("tf32_tcgen05_tflops") or (peaks.get("bf16_tflops") or 1620.5) / 2.0, \
            "tcgen05 kind::tf32" + ("" if peaks.get("tf32_tcgen05_tflops") else " (peak = half the measured bf16 figure; not measured directly)"), "cgemm_tc_kernel"
        issued, what = achieved * 3.0, "; 3 TF32 products issued per algorithmic product (hi.hi + hi.lo + lo.hi)"
    res = {"cfg": 5, "workload": f"power_method_D{D}_N{N}_K{K}_{cdt_name}", "metric": "transfer_matrix_applications_per_sec",
           "unit": "applications/s", "dtype": cdt_name, "value": apps / ms * 1e3, "ms_per_step": ms, "units_per_step": apps,
           "api": "qmps_tm_power",
           "roofline": _roof("tensor", issued, peak, "TOP/s" if i8 else "TFLOP/s", kern, f"{flops:.4g} real flops (32 D^3)" + what, note=pipe)}
    if i8:
        res["roofline"]["vs_fp64_tensor_pipe"] = achieved / peaks["fp64_dmma_tflops"]
    res["roofline"]["algorithmic_tflops"] = achieved
    if not with_e2e:
        return res
    Ah, Bh = A.cpu().pin_memory(), Bt.cpu().pin_memory()          # the caller's tensors in pinned host memory

    def e2e():
        r, ray = B.tm_power(Ah.to(dev, non_blocking=True), Bh.to(dev, non_blocking=True), K)
        return r.cpu().numpy(), ray.cpu().numpy()
    ms_e = wall_ms(e2e, reps=2, warm=1)
    res["e2e"] = {"value": apps / ms_e * 1e3, "unit": "applications/s", "h2d_bytes_per_step": Ah.numel() * Ah.element_size() + Bh.numel() * Bh.element_size(),
                  "d2h_bytes_per_step": N * D * D * Ah.element_size() + N * Ah.element_size(), "api": "batched.tm_power on host tensors (pinned H2D of A and B, D2H of r_K and the Rayleigh quotients)"}
    return res


def leg_scalar_latency(torch, dev):
    """Batch-of-one drop-in latency: ``qmps_b200.tools.get_env_exact(U)`` (one C-ABI call on host buffers, one
    synchronisation) against the oracle port of qmps/tools.py:176-182 on one host core, same 4 x 4 unitaries."""
    import time
    from scipy.stats import unitary_group
    from qmps_b200 import tools as T
    import oracle as O
    Us = [unitary_group.rvs(4, random_state=300 + k) for k in range(64)]
    for U in Us[:8]:
        T.get_env_exact(U); O.get_env_exact(U)
    t0 = time.perf_counter()
    for _ in range(4):
        for U in Us:
            T.get_env_exact(U)
    gpu_us = (time.perf_counter() - t0) / (4 * len(Us)) * 1e6
    t0 = time.perf_counter()
    for U in Us:
        O.get_env_exact(U)
    cpu_us = (time.perf_counter() - t0) / len(Us) * 1e6
    err = max(float(np.abs(np.abs(T.get_env_exact(U)[:, 0]) - np.abs(O.get_env_exact(U)[:, 0])).max()) for U in Us[:8])
    # the scalar cost function the reference's optimisers call (ground_state.py:150-168): one parameter vector per call
    from qmps_b200.ground_state import SparseFullEnergyOptimizer
    Hm = O.tfim_matrix(1.0)
    opt = SparseFullEnergyOptimizer(Hm, D=2, depth=2)
    ths = np.random.default_rng(5).normal(size=(64, len(opt.initial_guess)))
    for th in ths[:8]:
        opt.objective_function(th)
    t0 = time.perf_counter()
    for _ in range(4):
        for th in ths:
            opt.objective_function(th)
    obj_us = (time.perf_counter() - t0) / (4 * len(ths)) * 1e6
    ref_fn = lambda th: O.energy_transfer(O.unitary_to_tensor(O.shallow_cnot_state_tensor(2, th)), Hm)   # noqa: E731
    t0 = time.perf_counter()
    for th in ths:
        ref_fn(th)
    obj_cpu_us = (time.perf_counter() - t0) / len(ths) * 1e6
    obj_err = max(abs(opt.objective_function(th) - ref_fn(th)) for th in ths[:8])
    return {"cfg": 1, "workload": "get_env_exact_batch_of_one_D2_c128 (scalar drop-in latency)", "metric": "latency_per_call", "unit": "us",
            "dtype": "c128", "value": gpu_us, "higher_is_better": False, "api": "qmps_b200.tools.get_env_exact -> qmps_get_env_exact_host (H2D, 3 kernels, D2H, 1 sync)",
            "cpu_baseline": {"value": cpu_us, "unit": "us", "cores": 1, "kind": "port", "sample": "64 per-call oracle.get_env_exact (qmps/tools.py:176-182 restated)"},
            "max_abs_diff_first_column_moduli": err,
            "objective_function": {"value": obj_us, "unit": "us", "cpu_port_us": obj_cpu_us, "max_abs_diff": obj_err,
                                   "api": "SparseFullEnergyOptimizer.objective_function -> qmps_energy_theta_host (one call, one sync)"},
            "roofline": {"bound": "latency", "achieved": None, "peak": None, "unit": None, "frac": None, "kernel": "env_d2_stream_kernel + env2u_kernel",
                         "note": "three launches + two PCIe copies + one synchronisation; nothing here is throughput-bound"}}


def leg_tdvp_large(torch, B, dev, peaks, D=64, N=64):
    """(f)-3 at large bond dimension: TDVP tangent vectors of N left-canonical D = 64 tensors, every O(D^3) contraction on the
    tcgen05 kind::i8 kernels (batched.tdvp_tangent_large); unit = one tangent vector."""
    g = torch.Generator(device=dev).manual_seed(64)
    Z = torch.randn((N, 2 * D, D), dtype=torch.float64, device=dev, generator=g) + 1j * torch.randn((N, 2 * D, D), dtype=torch.float64, device=dev, generator=g)
    Q, _ = torch.linalg.qr(Z)
    A = Q.reshape(N, D, 2, D).permute(0, 2, 1, 3).contiguous()
    h = torch.tensor(np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0])) * -1.0 + 0.7 * 0.5 * (np.kron([[0, 1], [1, 0]], np.eye(2)) + np.kron(np.eye(2), [[0, 1], [1, 0]])),
                     dtype=torch.complex128, device=dev)
    info = {}

    def fn():
        out = B.tdvp_tangent_large(A, h)
        info.update(out[2])
        return out
    ms = timed_ms(torch, fn, reps=3, warm=2)
    d = 2
    gemms = 5 * d * d + 4 * d + 2 * d * info["k_iterations"] + 2 * d * (info["r_iterations"] + 1)
    flops = gemms * 8.0 * D ** 3
    ach = N * flops / ms * 1e3 / 1e12
    peak = peaks.get("i8_tcgen05_tops") or 2.0 * (peaks.get("bf16_tflops") or 1620.5)
    return {"cfg": 5, "workload": f"tdvp_tangent_D{D}_N{N}_c128 (classical iTDVP step on the tensor-core contraction)", "metric": "tdvp_tangents_per_sec",
            "unit": "tangents/s", "dtype": "c128", "value": N / ms * 1e3, "ms_per_step": ms, "units_per_step": N,
            "api": "batched.tdvp_tangent_large (host-level composition over qmps_tm_power / qmps_zgemm_c128_i8)",
            "iterations": {"power_method_r": info["r_iterations"], "neumann_K": info["k_iterations"]},
            "roofline": _roof("tensor", ach * 21.0, peak, "TOP/s", "zgemm_i8_kernel + slice_kernel",
                              f"{flops:.4g} real flops ({gemms} complex D^3 products at the measured iteration counts); 21 int8 products issued per real product",
                              note="tcgen05 kind::i8 (TOP/s); many small launches with host-side convergence checks: latency, not the tensor pipe, bounds it")}
