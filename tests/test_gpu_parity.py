"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same
seeded inputs, plus size-independent properties at full size.  Tolerances follow the
north star: 1e-10 relative in complex128, 1e-5 in complex64 (gap-aware where the
problem's condition number enters)."""
import numpy as np
import pytest
from scipy.stats import unitary_group

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def env():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from qmps_b200 import batched, represent, _lib
    _lib.require_device()
    import oracle as O
    return dict(torch=torch, B=batched, R=represent, O=O, L=_lib)


def haar_batch(n, count, seed):
    return np.stack([unitary_group.rvs(n, random_state=seed + k) for k in range(count)])


def tensors(D, count, seed, O):
    return np.ascontiguousarray(np.stack([O.unitary_to_tensor(u) for u in haar_batch(2 * D, count, seed)]))


def gap_of(A, O):
    w = np.sort(np.abs(np.linalg.eigvals(O.transfer_matrix(A))))[::-1]
    return 1.0 - w[1]


# ---------------------------------------------------------------- a4/a5, D = 2 fast path
@pytest.mark.parametrize("form", ["A", "U"])
@pytest.mark.parametrize("want_C", [True, False])
def test_env_d2_vs_oracle(env, form, want_C):
    t, B, O = env["torch"], env["B"], env["O"]
    N = 1000                                     # not a multiple of 32: ragged last tile
    U = haar_batch(4, N, 10)
    A = np.stack([O.unitary_to_tensor(u) for u in U])
    res = B.env_exact(A=t.from_numpy(A).cuda(), want_C=want_C) if form == "A" else B.env_exact(U=t.from_numpy(U).cuda(), want_C=want_C)
    eta, r = res.eta.cpu().numpy(), res.r.cpu().numpy()
    assert int(res.status.abs().sum()) == 0
    for k in range(N):
        e0, r0, C0, _ = O.env_exact_parts(A[k])
        tol = TOL / min(1.0, gap_of(A[k], O))
        assert abs(eta[k] - e0) < TOL
        assert np.abs(r[k] - r0).max() < tol
        if want_C:
            assert np.abs(res.C[k].cpu().numpy() - C0).max() < 10 * tol / np.sqrt(np.linalg.eigvalsh(r0)[0])


def test_env_d2_full_size_properties(env):
    """2^20 problems (BASELINE config 2): fixed-point equation, trace, hermiticity, r = C C^dagger."""
    t, B = env["torch"], env["B"]
    N = 1 << 20
    g = t.Generator(device="cuda").manual_seed(1)
    Z = t.randn((N, 4, 2), dtype=t.float64, device="cuda", generator=g) + 1j * t.randn((N, 4, 2), dtype=t.float64, device="cuda", generator=g)
    Q, _ = t.linalg.qr(Z)                                    # iso[(i,s), j], columns orthonormal
    A = Q.reshape(N, 2, 2, 2).permute(0, 2, 1, 3).contiguous()   # A[s,i,j] = iso[2i+s, j]
    res = B.env_exact(A=A)
    r, C, eta = res.r, res.C, res.eta
    ok = res.status == 0
    assert ok.float().mean() > 0.999
    phi = t.einsum("nsij,njl,nskl->nik", A, r, A.conj())
    gap_scale = 1e-9
    assert (phi - r)[ok].abs().max() < gap_scale
    assert (t.einsum("nii->n", r).real - 1).abs().max() < 1e-12
    assert (r - r.conj().transpose(1, 2)).abs().max() == 0
    assert (eta[ok] - 1).abs().max() < 1e-12
    assert (C @ C.conj().transpose(1, 2) - r)[ok].abs().max() < 1e-12
    assert C[:, 0, 1].abs().max() == 0 and (C[:, 0, 0].imag.abs().max() == 0)


@pytest.mark.parametrize("N,use_U", [(1, False), (1000, True), (4099, False)])
def test_env_d2_packed_records_equal_full_outputs(env, N, use_U):
    """qmps_env_exact_packed(+_host): the 64-byte records expand to the eta / r / C / status of qmps_env_exact."""
    t, B, O = env["torch"], env["B"], env["O"]
    from scipy.stats import unitary_group
    U = np.stack([unitary_group.rvs(4, random_state=7000 + k) for k in range(min(N, 64))])
    U = np.tile(U, (N // len(U) + 1, 1, 1))[:N].copy()
    A = np.stack([O.unitary_to_tensor(u) for u in U[:64]]); A = np.tile(A, (N // len(A) + 1, 1, 1, 1))[:N].copy()
    if N > 1:                                                # a product state: status NOT_PD travels in the record
        U[N // 2] = np.eye(4)
        A[N // 2] = O.unitary_to_tensor(np.eye(4))
    x = U if use_U else A
    full = B.env_exact(**({"U": t.from_numpy(x).cuda()} if use_U else {"A": t.from_numpy(x).cuda()}))
    packed = B.env_exact_packed(**({"U": t.from_numpy(x).cuda()} if use_U else {"A": t.from_numpy(x).cuda()})).cpu().numpy()
    host = B.env_exact_packed_host(x, is_unitary=use_U)
    assert np.array_equal(packed, host)
    eta, r, C, st = B.unpack_env(packed)
    st0 = full.status.cpu().numpy()
    assert np.array_equal(st, st0) and (N == 1 or st[N // 2] == env["L"].ST_NOT_PD)
    ok = st0 == 0
    assert np.abs(r[ok] - full.r.cpu().numpy()[ok]).max() < 1e-12
    assert np.abs(C[ok] - full.C.cpu().numpy()[ok]).max() < 1e-12
    assert np.abs(eta[ok] - full.eta.cpu().numpy()[ok]).max() < 1e-12


def test_env_d2_not_positive_definite_is_flagged(env):
    """Product state: r has rank one, the reference's cholesky raises LinAlgError (tools.py:182)."""
    t, B, L = env["torch"], env["B"], env["L"]
    U = np.eye(4, dtype=np.complex128)[None]
    res = B.env_exact(U=t.from_numpy(U).cuda())
    assert int(res.status[0]) == L.ST_NOT_PD
    from qmps_b200 import tools
    with pytest.raises(np.linalg.LinAlgError):
        tools.get_env_exact(U[0])


def test_env_empty_and_single(env):
    t, B = env["torch"], env["B"]
    res = B.env_exact(A=t.zeros((0, 2, 2, 2), dtype=t.complex128, device="cuda"))
    assert res.r.shape == (0, 2, 2)
    A = tensors(2, 1, 5, env["O"])
    res = B.env_exact(A=t.from_numpy(A).cuda())
    _, r0, _, _ = env["O"].env_exact_parts(A[0])
    assert np.abs(res.r[0].cpu().numpy() - r0).max() < TOL


def test_env_d2_complex64(env):
    t, B, O = env["torch"], env["B"], env["O"]
    A = tensors(2, 256, 77, O)
    res = B.env_exact(A=t.from_numpy(A).cuda().to(t.complex64))
    for k in range(256):
        _, r0, C0, _ = O.env_exact_parts(A[k])
        assert np.abs(res.r[k].cpu().numpy() - r0).max() < 1e-5 / min(1.0, gap_of(A[k], O))


# ---------------------------------------------------------------- a4/a5 generic D
@pytest.mark.parametrize("D,count", [(2, 64), (4, 64), (8, 24), (16, 3)])
@pytest.mark.parametrize("lc", [True, False])
def test_env_generic_vs_oracle(env, D, count, lc):
    t, B, O = env["torch"], env["B"], env["O"]
    if D == 2 and lc:
        pytest.skip("covered by the fast path tests")
    if D == 16 and not lc:
        count = 1
    A = tensors(D, count, 200 + D, O)
    res = B.env_exact(A=t.from_numpy(A).cuda(), assume_left_canonical=lc)
    assert int(res.status.abs().sum()) == 0
    for k in range(count):
        e0, r0, C0, _ = O.env_exact_parts(A[k])
        tol = 10 * TOL / min(1.0, gap_of(A[k], O))
        assert abs(res.eta[k].cpu().numpy() - e0) < 1e-9
        assert np.abs(res.r[k].cpu().numpy() - r0).max() < tol
        assert np.abs(res.C[k].cpu().numpy() - C0).max() < 100 * tol


def test_env_d8_tensor_pipe_kernel_edge_cases(env):
    """The D = 8 blocked elimination on the FP64 tensor pipe (env_dmma_kernel, the complex128 default) against the
    row-per-thread kernel it replaced and the oracle: Haar tensors, ansatz tensors at theta = 0 and at a small angle
    (environments with tiny Schmidt values), a product state (singular fixed-point system), d = 1, 3, 4 (the
    run-time physical-dimension build), and a batch that is not a multiple of anything."""
    t, B, O, L, R = env["torch"], env["B"], env["O"], env["L"], env["R"]
    lib = L.load()
    rng = np.random.default_rng(88)
    cases = []
    cases.append(tensors(8, 37, 808, O))
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(24)).program()
    th = np.concatenate([np.zeros((1, 24)), 1e-3 * rng.normal(size=(3, 24)), rng.normal(size=(9, 24))])
    cases.append(B.ansatz_tensors(prog, t.from_numpy(th).cuda()).cpu().numpy())
    for d in (1, 3, 4):                                       # left-canonical tensors with another physical dimension
        Z = rng.normal(size=(5, d * 8, 8)) + 1j * rng.normal(size=(5, d * 8, 8))
        if d == 1:
            Q = np.stack([np.linalg.qr(z)[0] for z in Z])     # unitary 8 x 8: r = 1/8 exactly
        else:
            Q = np.stack([np.linalg.qr(z)[0] for z in Z])
        cases.append(np.ascontiguousarray(Q.reshape(5, 8, d, 8).transpose(0, 2, 1, 3)))
    for A in cases:
        Ad = t.from_numpy(np.ascontiguousarray(A)).cuda()
        outs = {}
        for v in (3, 0):
            lib.qmps_set_option(b"er_wide", v)
            try:
                outs[v] = B.env_exact(A=Ad)
                t.cuda.synchronize()
            finally:
                lib.qmps_set_option(b"er_wide", -1)
        st3, st0 = outs[3].status.cpu().numpy(), outs[0].status.cpu().numpy()
        ok = (st3 == 0) & (st0 == 0)
        # the two kernels may disagree on the status only where the system is singular to rounding (d = 1: r is not unique)
        if A.shape[1] > 1:
            assert np.array_equal(st3 != 0, st0 != 0)
        r3, r0 = outs[3].r.cpu().numpy(), outs[0].r.cpu().numpy()
        for k in np.nonzero(ok)[0]:
            if A.shape[1] == 1:
                continue
            gap = min(1.0, gap_of(A[k], O))
            assert np.abs(r3[k] - r0[k]).max() < 10 * TOL / gap
            _, rr, CC, _ = O.env_exact_parts(A[k])
            assert np.abs(r3[k] - rr).max() < 10 * TOL / gap
            assert abs(np.trace(r3[k]) - 1) < 1e-12 and np.abs(r3[k] - r3[k].conj().T).max() == 0
            if np.linalg.eigvalsh(rr)[0] > 1e-9:
                assert np.abs(outs[3].C[k].cpu().numpy() - CC).max() < 1000 * TOL / gap / np.sqrt(np.linalg.eigvalsh(rr)[0])
    # a product state: Phi(r) = r has a degenerate solution space -> flagged, not a silent wrong answer
    Ap = np.zeros((1, 2, 8, 8), dtype=complex)
    Ap[0, 0] = np.eye(8)
    res = B.env_exact(A=t.from_numpy(Ap).cuda())
    assert int(res.status[0]) in (L.ST_SINGULAR, L.ST_NOT_PD)


def test_env_general_tensor_not_canonical(env):
    t, B, O = env["torch"], env["B"], env["O"]
    rng = np.random.default_rng(4)
    A = rng.normal(size=(32, 2, 4, 4)) + 1j * rng.normal(size=(32, 2, 4, 4))
    res = B.env_exact(A=t.from_numpy(A).cuda(), assume_left_canonical=False, want_C=False)
    for k in range(32):
        e0, _, r0 = O.eigs(A[k])
        assert abs(res.eta[k].cpu().numpy() - e0) < 1e-9 * abs(e0)
        assert np.abs(res.r[k].cpu().numpy() - r0).max() < 1e-9


def test_env_two_site_block(env):
    """d = 4 input (merge(A1, A2)) as used by NonSparseFullTwoSiteEnergyOptimizer.env_function."""
    t, B, O = env["torch"], env["B"], env["O"]
    A = tensors(2, 32, 31, O)
    M = np.stack([O.merge(A[k], A[(k + 1) % 32]) for k in range(32)])
    res = B.env_exact(A=t.from_numpy(M).cuda())
    for k in range(32):
        _, r0, C0, _ = O.env_exact_parts(M[k])
        assert np.abs(res.r[k].cpu().numpy() - r0).max() < 1e-9


# ---------------------------------------------------------------- a6/a8/a11 mixed fixed points
@pytest.mark.parametrize("D,count", [(2, 128), (4, 48), (8, 6)])
@pytest.mark.parametrize("left", [False, True])
def test_fixed_point_vs_oracle(env, D, count, left):
    t, B, O = env["torch"], env["B"], env["O"]
    A, Bt = tensors(D, count, 300 + D, O), tensors(D, count, 900 + D, O)
    fp = B.fixed_point(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), left=left)
    vec_tr = B.fixed_point(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), left=left, gauge="trace").vec.cpu().numpy()
    assert int(fp.status.abs().sum()) == 0
    eta, vec = fp.eta.cpu().numpy(), fp.vec.cpu().numpy()
    for k in range(count):
        x0, v0 = (O.left_fixed_point if left else O.right_fixed_point)(A[k], Bt[k])
        # default gauge = zgeev's (the notebook-recorded xmps convention): largest entry real positive
        big = vec[k].reshape(-1)[np.argmax(np.abs(vec[k]))]
        assert abs(big.imag) < 1e-12 and big.real > 0
        assert abs(np.trace(vec_tr[k]).imag) < 1e-12 and np.trace(vec_tr[k]).real >= 0
        assert abs(abs(np.vdot(vec_tr[k], vec[k])) - 1) < 1e-10        # same ray, different phase
        E = O.transfer_matrix(A[k], Bt[k])
        Em = E.conj().T if left else E
        w = np.sort(np.abs(np.linalg.eigvals(E)))[::-1]
        gap = max(w[0] - w[1], 1e-3)
        assert abs(abs(eta[k]) - abs(x0)) < TOL * 10
        assert abs(fp.fid[k].item() - abs(x0) ** 2) < TOL * 10
        assert abs(fp.cost[k].item() + np.sqrt(abs(x0))) < TOL * 10
        assert abs(fp.echo[k].item() + np.log(abs(x0) ** 2)) < TOL * 100
        v = vec[k].reshape(-1)
        assert abs(np.linalg.norm(v) - 1) < 1e-12
        assert np.abs(Em @ v - eta[k] * v).max() < 1e-11 / gap
        if abs(eta[k] - x0) < 1e-9:                 # same eigenvalue picked: same gauge-fixed vector
            assert np.abs(vec[k] - v0).max() < 1e-9 / gap


@pytest.mark.parametrize("left", [False, True])
@pytest.mark.parametrize("d", [2, 4])
def test_fixed_point_d4_eigenvalue_only_kernel(env, left, d):
    """The register-resident D = 4 kernel (kernels_fp16.cuh, eigenvalue only) against the oracle's
    dense eig and against the generic shared-memory kernel on the same inputs; d = 4 is the
    two-site (merged) map of the Loschmidt cost.  An odd batch size leaves half a warp idle."""
    t, B, O, L = env["torch"], env["B"], env["O"], env["L"]
    count = 101
    A, Bt = tensors(4, count, 1300 + d, O), tensors(4, count, 1900 + d, O)
    if d == 4:
        A = np.ascontiguousarray(np.einsum("naik,nbkj->nabij", A, A[::-1]).reshape(count, 4, 4, 4))
        Bt = np.ascontiguousarray(np.einsum("naik,nbkj->nabij", Bt, Bt[::-1]).reshape(count, 4, 4, 4))
    Ad, Bd = t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda()
    lib = L.load()
    default = B.fixed_point(Ad, Bd, left=left, want_vec=False)   # default: shared-resident quarter-warp kernel
    lib.qmps_set_option(b"fp16_fast", 0)
    try:
        slow = B.fixed_point(Ad, Bd, left=left, want_vec=False)  # generic shared-memory group kernel
    finally:
        lib.qmps_set_option(b"fp16_fast", 8)
    assert (default.eta.abs() - slow.eta.abs()).abs().max().item() < TOL
    for variant in (1, 3, 6, 7, 8, 9, 10):  # half-warp / quarter-warp register forms, shared-resident half / quarter warp (8: branch-free rsqrt, 9: packed two-kernel form, 10: trimmed sweep bodies)
        lib.qmps_set_option(b"fp16_fast", variant)
        try:
            fast = B.fixed_point(Ad, Bd, left=left, want_vec=False)
            f32 = B.fixed_point(Ad.to(t.complex64), Bd.to(t.complex64), left=left, want_vec=False)
            t.cuda.synchronize()
        finally:
            lib.qmps_set_option(b"fp16_fast", 8)
        assert int(fast.status.abs().sum()) == 0
        assert (fast.eta.abs() - slow.eta.abs()).abs().max().item() < TOL
        for k in range(count):
            x0 = (O.left_fixed_point if left else O.right_fixed_point)(A[k], Bt[k])[0]
            assert abs(abs(fast.eta[k].item()) - abs(x0)) < TOL
            assert abs(fast.cost[k].item() + np.sqrt(abs(x0))) < TOL
            w = np.linalg.eigvals(O.transfer_matrix(A[k], Bt[k]))
            lam = np.conj(fast.eta[k].item()) if left else fast.eta[k].item()
            assert np.abs(w - lam).min() < 1e-11               # it IS an eigenvalue of E, not just the right modulus
        assert (f32.eta.abs().double() - fast.eta.abs()).abs().max().item() < 1e-5      # complex64 mode


@pytest.mark.parametrize("left", [False, True])
@pytest.mark.parametrize("d", [2, 4])
def test_fixed_point_d2_register_kernel(env, left, d):
    """The thread-per-problem D = 2 kernel (kernels_fpd2.cuh, eigenvalue only; the default when no
    eigenvector is requested) against the oracle's dense eig and against the generic shared-memory
    kernel on the same inputs; d = 4 is the merged two-site map of the Loschmidt cost.  Includes
    identical pairs (eta = 1), near-identical pairs, the outer-product mode and an odd batch."""
    t, B, O, L = env["torch"], env["B"], env["O"], env["L"]
    count = 333
    A, Bt = tensors(2, count, 2300 + d, O), tensors(2, count, 2900 + d, O)
    Bt[:10] = A[:10]
    Bt[10:20] = A[10:20] + 1e-4 * np.random.default_rng(1).normal(size=(10, 2, 2, 2))
    if d == 4:
        A = np.ascontiguousarray(np.einsum("naik,nbkj->nabij", A, A).reshape(count, 4, 2, 2))
        Bt = np.ascontiguousarray(np.einsum("naik,nbkj->nabij", Bt, Bt).reshape(count, 4, 2, 2))
    Ad, Bd = t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda()
    lib = L.load()
    fast = B.fixed_point(Ad, Bd, left=left, want_vec=False)
    f32 = B.fixed_point(Ad.to(t.complex64), Bd.to(t.complex64), left=left, want_vec=False)
    outer = B.fixed_point(Ad[:7], Bd[:9], pair="outer", left=left, want_vec=False)
    lib.qmps_set_option(b"fp_d2", 0)
    try:
        slow = B.fixed_point(Ad, Bd, left=left, want_vec=False)
        t.cuda.synchronize()
    finally:
        lib.qmps_set_option(b"fp_d2", 1)
    assert int(fast.status.abs().sum()) == 0
    assert (fast.eta.abs() - slow.eta.abs()).abs().max().item() < TOL
    assert (fast.cost - slow.cost).abs().max().item() < TOL and (fast.echo - slow.echo).abs().max().item() < TOL * 10
    assert np.abs(np.abs(fast.eta[:10].cpu().numpy()) - 1).max() < 1e-12
    for k in range(0, count, 3):
        x0 = (O.left_fixed_point if left else O.right_fixed_point)(A[k], Bt[k])[0]
        assert abs(abs(fast.eta[k].item()) - abs(x0)) < TOL
        assert abs(fast.fid[k].item() - abs(x0) ** 2) < TOL
        w = np.linalg.eigvals(O.transfer_matrix(A[k], Bt[k]))
        lam = np.conj(fast.eta[k].item()) if left else fast.eta[k].item()
        assert np.abs(w - lam).min() < 1e-11
    for i in range(7):
        for j in range(9):
            x0 = (O.left_fixed_point if left else O.right_fixed_point)(A[i], Bt[j])[0]
            assert abs(abs(outer.eta[i, j].item()) - abs(x0)) < TOL
    assert (f32.eta.abs().double() - fast.eta.abs()).abs().max().item() < 1e-5      # complex64 mode


def test_fixed_point_d4_degenerate_inputs(env):
    """Edge cases of the QR kernel: identical tensors (eta = 1 exactly known), a product state
    (rank-one map: 15 zero eigenvalues) and the zero tensor."""
    t, B, O = env["torch"], env["B"], env["O"]
    A = tensors(4, 4, 5, O)
    prod = np.zeros((1, 2, 4, 4), complex); prod[0, 0, 0, 0] = 1.0
    zero = np.zeros((1, 2, 4, 4), complex)
    X = t.from_numpy(np.concatenate([A, prod, zero])).cuda()
    lib = env["L"].load()
    for fast in (0, 1, 3, 6, 7, 8, 9, 10):
        lib.qmps_set_option(b"fp16_fast", fast)
        try:
            fp = B.fixed_point(X, X, want_vec=False)
            eta = fp.eta.cpu().numpy()
        finally:
            lib.qmps_set_option(b"fp16_fast", 8)
        assert np.abs(np.abs(eta[:4]) - 1).max() < 1e-12
        assert abs(abs(eta[4]) - 1) < 1e-12 and abs(eta[5]) == 0


@pytest.mark.parametrize("left", [False, True])
@pytest.mark.parametrize("d", [2, 4])
def test_fixed_point_d8_warp_kernel(env, left, d):
    """The warp-per-problem D = 8 kernels (kernels_fp64p.cuh, the default, and kernels_fp64w.cuh) against numpy's dense eig
    of the oracle's transfer matrix and against the generic CTA-per-problem kernel on the same inputs; d = 4 is the
    two-site (merged) map of the Loschmidt cost.  37 problems: fewer than the persistent grid, an odd count."""
    t, B, O, L = env["torch"], env["B"], env["O"], env["L"]
    count = 37
    A, Bt = tensors(8, count, 2300 + d, O), tensors(8, count, 2900 + d, O)
    if d == 4:
        A = np.ascontiguousarray(np.einsum("naik,nbkj->nabij", A, A[::-1]).reshape(count, 4, 8, 8))
        Bt = np.ascontiguousarray(np.einsum("naik,nbkj->nabij", Bt, Bt[::-1]).reshape(count, 4, 8, 8))
    Ad, Bd = t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda()
    lib = L.load()
    fast = B.fixed_point(Ad, Bd, left=left, want_vec=False)     # default: packed two-kernel form (kernels_fp64p.cuh)
    f32 = B.fixed_point(Ad.to(t.complex64), Bd.to(t.complex64), left=left, want_vec=False)
    try:
        lib.qmps_set_option(b"fp64_fast", 0)
        slow = B.fixed_point(Ad, Bd, left=left, want_vec=False)  # generic CTA-per-problem kernel
        lib.qmps_set_option(b"fp64_fast", 1)
        one = B.fixed_point(Ad, Bd, left=left, want_vec=False)   # one-kernel warp-per-problem form (kernels_fp64w.cuh)
        one32 = B.fixed_point(Ad.to(t.complex64), Bd.to(t.complex64), left=left, want_vec=False)
    finally:
        lib.qmps_set_option(b"fp64_fast", 2)
    assert int(fast.status.abs().sum()) == 0 and int(one.status.abs().sum()) == 0
    assert (fast.eta - slow.eta).abs().max().item() < 1e-12
    assert (one.eta - slow.eta).abs().max().item() < 1e-12
    assert (one32.eta.abs().double() - fast.eta.abs()).abs().max().item() < 1e-5
    for k in range(count):
        E = sum(np.kron(A[k, s], Bt[k, s].conj()) for s in range(d))
        w = np.linalg.eigvals(E)
        x0 = np.abs(w).max()
        assert abs(abs(fast.eta[k].item()) - x0) < TOL
        assert abs(fast.cost[k].item() + np.sqrt(x0)) < TOL
        assert abs(fast.fid[k].item() - x0 ** 2) < TOL
        assert abs(fast.echo[k].item() + np.log(x0 ** 2)) < TOL * 10
        lam = np.conj(fast.eta[k].item()) if left else fast.eta[k].item()
        assert np.abs(w - lam).min() < 1e-11                   # it IS an eigenvalue of E, not just the right modulus
    assert (f32.eta.abs().double() - fast.eta.abs()).abs().max().item() < 1e-5          # complex64 mode
    # outer pairing (the Loschmidt grid: one reference tensor against many) goes through the same kernel
    outer = B.fixed_point(Ad[:2], Bd[:5], pair="outer", left=left, want_vec=False)
    for i in range(2):
        for j in range(5):
            E = sum(np.kron(A[i, s], Bt[j, s].conj()) for s in range(d))
            assert abs(abs(outer.eta[i, j].item()) - np.abs(np.linalg.eigvals(E)).max()) < TOL


def test_fixed_point_d8_degenerate_inputs(env):
    """Edge cases of the D = 8 QR kernel: identical tensors (|eta| = 1), a product state (rank-one map: 63 zero
    eigenvalues, every sub-diagonal entry deflates at once), the zero tensor, and a block-diagonal pair whose 64 x 64 map
    splits in the middle of the matrix (windows that start above row 32)."""
    t, B, O = env["torch"], env["B"], env["O"]
    A = tensors(8, 3, 15, O)
    prod = np.zeros((1, 2, 8, 8), complex); prod[0, 0, 0, 0] = 1.0
    zero = np.zeros((1, 2, 8, 8), complex)
    blk = np.zeros((1, 2, 8, 8), complex)
    A4 = tensors(4, 2, 16, O)
    blk[0, :, :4, :4] = A4[0]; blk[0, :, 4:, 4:] = 0.5 * A4[1]
    X = t.from_numpy(np.concatenate([A, prod, zero, blk])).cuda()
    lib = env["L"].load()
    res = {}
    for fast in (0, 1, 2):
        lib.qmps_set_option(b"fp64_fast", fast)
        try:
            fp = B.fixed_point(X, X, want_vec=False)
            res[fast] = fp.eta.cpu().numpy()
            assert int(fp.status.abs().sum().item()) == 0
        finally:
            lib.qmps_set_option(b"fp64_fast", 2)
        eta = res[fast]
        assert np.abs(np.abs(eta[:3]) - 1).max() < 1e-12
        assert abs(abs(eta[3]) - 1) < 1e-12 and abs(eta[4]) == 0
        assert abs(abs(eta[5]) - 1) < 1e-12                    # the leading block is an isometry with itself
    assert np.abs(np.abs(res[0]) - np.abs(res[1])).max() < 1e-12 and np.abs(np.abs(res[0]) - np.abs(res[2])).max() < 1e-12


def test_fixed_point_packed_forms_chunked_workspace(env):
    """The packed two-kernel forms work on slices of the batch (one workspace of at most 32 768 problems at D = 8, 2^20 at
    D = 4): batches larger than one slice must give the same eigenvalues as the one-kernel forms, problem by problem
    (outer pairing: problem index -> (ia, ib) goes through the slice offset)."""
    t, B, O, L = env["torch"], env["B"], env["O"], env["L"]
    lib = L.load()
    A8, B8 = t.from_numpy(tensors(8, 190, 41, O)).cuda(), t.from_numpy(tensors(8, 180, 42, O)).cuda()      # 34 200 problems
    try:
        lib.qmps_set_option(b"fp64_fast", 2)
        packed = B.fixed_point(A8, B8, pair="outer", want_vec=False)
        lib.qmps_set_option(b"fp64_fast", 1)
        one = B.fixed_point(A8, B8, pair="outer", want_vec=False)
    finally:
        lib.qmps_set_option(b"fp64_fast", 2)
    assert int(packed.status.abs().sum()) == 0
    assert (packed.eta - one.eta).abs().max().item() < 1e-12
    assert (packed.cost - one.cost).abs().max().item() < 1e-12
    A4, B4 = t.from_numpy(tensors(4, 1030, 43, O)).cuda(), t.from_numpy(tensors(4, 1020, 44, O)).cuda()    # 1 050 600 problems
    try:
        lib.qmps_set_option(b"fp16_fast", 9)
        packed = B.fixed_point(A4, B4, pair="outer", want_vec=False)
        lib.qmps_set_option(b"fp16_fast", 8)
        one = B.fixed_point(A4, B4, pair="outer", want_vec=False)
    finally:
        lib.qmps_set_option(b"fp16_fast", 8)
    assert int(packed.status.abs().sum()) == 0
    assert (packed.eta - one.eta).abs().max().item() < 1e-12


def test_fixed_point_outer_and_broadcast(env):
    t, B, O = env["torch"], env["B"], env["O"]
    A, Bt = tensors(2, 3, 1, O), tensors(2, 5, 2, O)
    fp = B.fixed_point(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), pair="outer")
    assert fp.eta.shape == (3, 5)
    for i in range(3):
        for j in range(5):
            assert abs(abs(fp.eta[i, j].item()) - abs(O.right_fixed_point(A[i], Bt[j])[0])) < TOL * 10
    fp = B.fixed_point(t.from_numpy(A[:1]).cuda(), t.from_numpy(Bt).cuda())
    for j in range(5):
        assert abs(abs(fp.eta[j].item()) - abs(O.right_fixed_point(A[0], Bt[j])[0])) < TOL * 10


def test_self_overlap_is_one(env):
    t, B, O = env["torch"], env["B"], env["O"]
    A = tensors(4, 16, 3, O)
    fp = B.fixed_point(t.from_numpy(A).cuda(), t.from_numpy(A).cuda())
    assert (fp.fid - 1).abs().max() < 1e-12


# ---------------------------------------------------------------- a1/a2/a3/a7
def test_unitary_to_tensor_and_back(env, golden):
    t, B = env["torch"], env["B"]
    g = golden["ref_tools"]
    for D in (2, 4, 8):
        U, Aref = g[f"u2t_U_D{D}"], g[f"u2t_A_D{D}"]
        A = B.unitary_to_tensor(t.from_numpy(U).cuda())
        assert np.array_equal(A.cpu().numpy(), Aref)          # pure index shuffle: bit exact
        U2 = B.tensor_to_unitary(A).cpu().numpy()
        n = 2 * D
        for k in range(len(U)):
            assert np.array_equal(U2[k][:, :D], g[f"t2u_U_D{D}"][k][:, :D])   # unique columns, exact
            assert np.abs(U2[k].conj().T @ U2[k] - np.eye(n)).max() < 1e-13
            assert np.array_equal(B.unitary_to_tensor(t.from_numpy(U2[k:k + 1]).cuda()).cpu().numpy()[0], Aref[k])


def test_environment_to_unitary(env, golden):
    t, B = env["torch"], env["B"]
    g = golden["ref_tools"]
    for key in ("e2u_in", "e2u_in_D4"):
        C = g[key]
        V = B.environment_to_unitary(t.from_numpy(C.reshape(1, -1)).cuda()).cpu().numpy()[0]
        ref = g[key.replace("in", "out")]
        n = V.shape[0]
        assert np.abs(V[:, 0] - ref[:, 0]).max() < 1e-15       # the unique column
        assert np.abs(V.conj().T @ V - np.eye(n)).max() < 1e-13


def test_merge(env):
    t, B, O = env["torch"], env["B"], env["O"]
    for D in (2, 4, 8):
        A, Bt = tensors(D, 6, 40 + D, O), tensors(D, 6, 60 + D, O)
        M = B.merge(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda()).cpu().numpy()
        for k in range(6):
            assert np.abs(M[k] - O.merge(A[k], Bt[k])).max() < 1e-14
    W = np.stack([O.tfim_evolution_gate(0.2, 0.1 * k) for k in range(5)])
    A = tensors(2, 1, 9, O)
    M = B.merge(t.from_numpy(A).cuda(), t.from_numpy(A).cuda(), t.from_numpy(W).cuda()).cpu().numpy()
    for k in range(5):
        assert np.abs(M[k] - O.apply_two_site_gate(W[k], O.merge(A[0], A[0]))).max() < 1e-14


# ---------------------------------------------------------------- a14 ansaetze
def ansatz_cases(R, O, rng):
    return [
        (R.ShallowFullStateTensor(2, rng.normal(size=15)), O.shallow_full_state_tensor),
        (R.ShallowCNOTStateTensor(2, rng.normal(size=4)), lambda p: O.shallow_cnot_state_tensor(2, p)),
        (R.ShallowCNOTStateTensor(8, rng.normal(size=6)), lambda p: O.shallow_cnot_state_tensor(8, p)),
        (R.ShallowCNOTStateTensor_nonuniform(4, rng.normal(size=12)), lambda p: O.shallow_cnot_state_tensor_nonuniform(4, p)),
        (R.ShallowCNOTStateTensor_nonuniform(8, rng.normal(size=24)), lambda p: O.shallow_cnot_state_tensor_nonuniform(8, p)),
        (R.ShallowCNOTStateTensor_nonuniform(16, rng.normal(size=10)), lambda p: O.shallow_cnot_state_tensor_nonuniform(16, p)),
        (R.ShallowCNOTStateTensor3(4, rng.normal(size=6)), lambda p: O.shallow_cnot_state_tensor3(4, p)),
        (R.ShallowQAOAStateTensor(4, rng.normal(size=6)), lambda p: O.shallow_qaoa_state_tensor(4, p)),
        (R.ExactAfter4(2, rng.normal(size=12)), lambda p: O.exact_after4(2, p)),
        (R.ExactAfter4(4, rng.normal(size=12)), lambda p: O.exact_after4(4, p)),
        (R.StateGate(rng.normal(size=6)), O.state_gate),
        (R.GSFAnsatz(rng.normal(size=8)), O.gsf_ansatz),
    ]


def test_ansatz_families(env):
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    rng = np.random.default_rng(3)
    for gate, ofn in ansatz_cases(R, O, rng):
        theta = rng.normal(size=(5, gate.p))
        U = B.ansatz_tensors(gate.program(), theta, full_unitary=True).cpu().numpy()
        A = B.ansatz_tensors(gate.program(), theta).cpu().numpy()
        for k in range(5):
            Uo = ofn(theta[k])
            assert np.abs(U[k] - Uo).max() < 1e-13, type(gate).__name__
            assert np.abs(A[k] - O.unitary_to_tensor(Uo)).max() < 1e-13
        assert np.abs(R.unitary(gate) - ofn(gate.params)).max() < 1e-13


# ---------------------------------------------------------------- a9/a12 energies and rotosolve
@pytest.mark.parametrize("D,layers,count", [(2, 3, 64), (4, 2, 32), (8, 3, 8)])
def test_energy_theta_with_shifts(env, D, layers, count):
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    nq = int(np.log2(D)) + 1
    rng = np.random.default_rng(50 + D)
    theta = rng.normal(size=(count, 2 * nq * layers))
    gate = R.ShallowCNOTStateTensor_nonuniform(D, theta[0])
    H = O.heisenberg_matrix() if D == 8 else O.tfim_matrix(1.0)
    coord = 3
    for shifts in (None, B.ROTO3_SHIFTS, B.ROTO6_SHIFTS):
        e = B.energy_theta(gate.program(), theta, H, coord=coord if shifts else None, shifts=shifts).cpu().numpy()
        for n in range(min(count, 6)):
            for s, sh in enumerate(shifts or (0.0,)):
                th = theta[n].copy(); th[coord] += sh
                ref = O.energy_transfer(O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(D, th)), H)
                got = e[n, s] if shifts else e[n]
                assert abs(got - ref) < 1e-9, (D, n, s, got, ref)


def test_energy_sfst_matches_statevector_route(env):
    """Against the reference's own route: get_env_exact -> State(U,V,2) -> simulate (ground_state.py:251-266)."""
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    rng = np.random.default_rng(8)
    theta = rng.normal(size=(16, 15))
    H = O.tfim_matrix(0.7)
    e = B.energy_theta(R.ShallowFullStateTensor(2, theta[0]).program(), theta, H).cpu().numpy()
    for n in range(16):
        assert abs(e[n] - O.energy_of_unitary(O.shallow_full_state_tensor(theta[n]), H)) < 1e-9


def test_energy_tensor_and_two_site(env):
    t, B, O = env["torch"], env["B"], env["O"]
    H = O.tfim_matrix(1.3)
    for D in (2, 4, 8):
        A = tensors(D, 8, 70 + D, O)
        e = B.energy_tensor(t.from_numpy(A).cuda(), H).cpu().numpy()
        for k in range(8):
            assert abs(e[k] - O.energy_transfer(A[k], H)) < 1e-9
    A = tensors(2, 8, 5, O)
    M = np.stack([O.merge(A[k], A[k + 1]) for k in range(0, 8, 2)] + [O.merge(A[k + 1], A[k]) for k in range(0, 8, 2)])
    e = B.energy_tensor(t.from_numpy(M).cuda(), H, two_site=True).cpu().numpy()
    for k in range(4):
        ref = O.energy_two_site_transfer(A[2 * k], A[2 * k + 1], H)
        assert abs(0.5 * (e[k] + e[k + 4]) - ref) < 1e-9
    U1, U2 = haar_batch(4, 2, 123)
    assert abs(O.energy_two_site_statevector(U1, U2, H) - O.energy_two_site_transfer(O.unitary_to_tensor(U1), O.unitary_to_tensor(U2), H)) < 1e-12


def test_rotosolve_fit_and_sweeps(env):
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    rng = np.random.default_rng(12)
    c3 = rng.normal(size=(100, 3))
    fit = B.rotosolve_fit(c3)
    ref = np.array([O.rotosolve_theta3(*row) for row in c3])
    assert np.abs(fit.theta_star.cpu().numpy() - ref).max() < 1e-13
    c6 = rng.normal(size=(100, 6))
    fit = B.rotosolve_fit(c6)
    ref = np.array([O.double_rotosolve_fit(*row) for row in c6])
    assert np.abs(fit.fit.cpu().numpy() - ref).max() < 1e-13
    ts = fit.theta_star.cpu().numpy()
    xs = np.linspace(-np.pi, np.pi, 20001)
    for k in range(100):
        a, b, c, d, P, u, Q, v = ref[k]
        f = lambda x: P * np.sin(2 * x + u) + Q * np.sin(x + v)
        assert f(ts[k]) <= f(xs).min() + 1e-9
    # a device sweep never increases the energy and matches a host replay of the same rule
    theta = rng.normal(size=(8, 12))
    gate = R.ShallowCNOTStateTensor_nonuniform(4, theta[0])
    H = O.tfim_matrix(1.0)
    th = t.from_numpy(theta.copy()).cuda()
    e0 = B.energy_theta(gate.program(), th, H).cpu().numpy()
    e1 = B.rotosolve_sweeps(gate.program(), th, H, n_sweeps=1).cpu().numpy()
    assert (e1 <= e0 + 1e-10).all()
    host = theta[0].copy()
    ef = lambda p: O.energy_transfer(O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(4, p)), H)
    for i in range(12):
        ei = np.eye(12)[i]
        host[i] = O.rotosolve_step3(host[i], ef(host), ef(host + ei * np.pi / 2), ef(host - ei * np.pi / 2))
    # coordinate 0 is rz on the |0> input qubit: a global phase, the cost is flat in it and the
    # closed-form angle is atan2(noise, noise); every other coordinate must agree
    assert np.abs(th[0].cpu().numpy() - host)[1:].max() < 1e-8
    assert abs(e1[0] - ef(host)) < 1e-9


# ---------------------------------------------------------------- a11 Loschmidt pipeline
def test_loschmidt_costs_d2_against_reference_circuit(env):
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    rng = np.random.default_rng(21)
    theta = rng.normal(size=(6, 15))
    A0 = O.unitary_to_tensor(O.shallow_full_state_tensor(rng.normal(size=15)))
    W = np.stack([O.tfim_evolution_gate(0.2, 0.04 * k) for k in range(5)])
    cost, echo, eta = B.loschmidt_costs(R.ShallowFullStateTensor(2, theta[0]).program(), theta, A0, W)
    cost, echo = cost.cpu().numpy(), echo.cpu().numpy()
    for p in range(6):
        Bp = O.unitary_to_tensor(O.shallow_full_state_tensor(theta[p]))
        for k in range(5):
            assert abs(cost[p, k] - O.loschmidt_cost(A0, Bp, W[k])) < TOL * 10
            assert abs(cost[p, k] - O.loschmidt_cost_circuit(A0, Bp, W[k])) < 1e-9      # the 6-qubit circuit
            assert abs(echo[p, k] - 4 * (-np.log(-cost[p, k]))) < 1e-9


def test_loschmidt_costs_d4(env):
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    rng = np.random.default_rng(2)
    theta = rng.normal(size=(16, 12))
    gate = R.ShallowCNOTStateTensor_nonuniform(4, theta[0])
    A0 = O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(4, theta[0]))
    W = np.stack([O.tfim_evolution_gate(0.2, 0.02 * k) for k in range(0, 1000, 97)])
    cost, echo, _ = B.loschmidt_costs(gate.program(), theta, A0, W)
    cost = cost.cpu().numpy()
    assert abs(cost[0, 0] + 1) < 1e-12                      # k = 0, same state: overlap 1
    for p in range(0, 16, 3):
        Bp = O.unitary_to_tensor(O.shallow_cnot_state_tensor_nonuniform(4, theta[p]))
        for k in range(len(W)):
            assert abs(cost[p, k] - O.loschmidt_cost(A0, Bp, W[k])) < TOL * 10


def test_loschmidt_rate_vs_reference_golden(env, golden):
    B = env["B"]
    g = golden["ref_exact_loschmidt"]
    got = B.loschmidt_rate(g["t"], 1.5, 0.2).cpu().numpy()
    assert np.abs(got - g["g15_02"]).max() < 1e-8
    got = B.loschmidt_rate(g["t"], 0.5, 2.0).cpu().numpy()
    assert np.abs(got - g["g05_20"]).max() < 1e-6          # crosses DQPT cusps: log singularities


# ---------------------------------------------------------------- cfg 5, (e)
@pytest.mark.parametrize("D", [16, 64, 100, 256])
def test_tm_power_vs_oracle(env, D):
    """cfg 5: D = 100 exercises the ragged edges of the 64 x 64 DMMA tiles, 256 the multi-tile norm."""
    t, B, O = env["torch"], env["B"], env["O"]
    cnt = 3 if D <= 100 else 2
    A, Bt = tensors(D, cnt, 400 + D, O), tensors(D, cnt, 500 + D, O)
    r, ray = B.tm_power(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), K=8)
    for k in range(cnt):
        r0, q0 = O.power_method(A[k], Bt[k], 8)
        assert np.abs(r[k].cpu().numpy() - r0).max() < 1e-12
        assert abs(ray[k].item() - q0) < 1e-12


@pytest.mark.parametrize("D", [64, 128])
def test_overlap_power_large_D_vs_sparse_eigensolver(env, D):
    """Large-D Loschmidt-echo / overlap step: the leading eigenvalue of E_AB from the power method on the tensor-core
    contraction against ARPACK on the same map applied as a linear operator (a dense D^2 x D^2 eig is out of reach)."""
    from scipy.sparse.linalg import LinearOperator, eigs
    t, B, O = env["torch"], env["B"], env["O"]
    cnt = 3
    A = tensors(D, cnt, 9100 + D, O)
    rng = np.random.default_rng(D)
    # B = A moved by a small unitary rotation of the physical index and a perturbation: a well-gapped, complex eta
    Bt = np.empty_like(A)
    for k in range(cnt):
        Z = A[k].transpose(1, 0, 2).reshape(2 * D, D) + 0.01 * (rng.normal(size=(2 * D, D)) + 1j * rng.normal(size=(2 * D, D))) / np.sqrt(D)
        Uz, _, Vz = np.linalg.svd(Z, full_matrices=False)       # closest isometry (polar factor): stays in A's gauge
        Q = (Uz @ Vz) * np.exp(0.3j)
        Bt[k] = Q.reshape(D, 2, D).transpose(1, 0, 2)
    res = B.overlap_power(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), tol=1e-10)
    assert bool(res.converged.all()) and res.iterations <= 2048
    eta = res.eta.cpu().numpy()
    for k in range(cnt):
        Bh = Bt[k].conj().transpose(0, 2, 1)
        op = LinearOperator((D * D, D * D), dtype=complex,
                            matvec=lambda v: np.sum(A[k] @ v.reshape(D, D) @ Bh, axis=0).reshape(-1))
        ref = eigs(op, k=1, which="LM", tol=1e-13, v0=np.eye(D).reshape(-1).astype(complex), maxiter=20000)[0][0]
        assert abs(eta[k] - ref) < 1e-8 * abs(ref)
        rk = res.r[k].cpu().numpy()
        assert np.abs(np.sum(A[k] @ rk @ Bh, axis=0) - eta[k] * rk).max() < 1e-6       # eigen-equation residual
    fid = B.overlap(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda()).cpu().numpy()      # dispatches to the power method for D > 16
    assert np.abs(fid - np.abs(eta) ** 2).max() < 1e-9
    assert np.abs(res.rate.cpu().numpy() + np.log(np.abs(eta) ** 2)).max() < 1e-12


@pytest.mark.parametrize("D,dt_name", [(8, "complex128"), (64, "complex128"), (100, "complex128"), (64, "complex64")])
def test_tm_apply_vs_numpy(env, D, dt_name):
    """qmps_tm_apply: one unnormalised application Y = sum_s A_s X B_s^dagger (general X, not Hermitian)."""
    t, B, O = env["torch"], env["B"], env["O"]
    rng = np.random.default_rng(3 * D)
    A, Bt = tensors(D, 3, 4400 + D, O), tensors(D, 3, 4500 + D, O)
    X = rng.normal(size=(3, D, D)) + 1j * rng.normal(size=(3, D, D))
    cdt = getattr(t, dt_name)
    Y = B.tm_apply(t.from_numpy(A).cuda().to(cdt), t.from_numpy(Bt).cuda().to(cdt), t.from_numpy(X).cuda().to(cdt)).cpu().numpy()
    ref = np.einsum("nsij,njk,nslk->nil", A, X, Bt.conj())
    tol = 1e-12 if dt_name == "complex128" else 2e-5
    assert np.abs(Y - ref).max() < tol * np.abs(ref).max()


def test_tm_power_same_tensor_converges_to_environment(env):
    """Size-independent property: for A = B left-canonical the power method converges to the
    exact environment of the direct solver (trace-normalised), Rayleigh quotient -> 1."""
    t, B, O = env["torch"], env["B"], env["O"]
    A = t.from_numpy(tensors(16, 4, 77, O)).cuda()
    r, ray = B.tm_power(A, A, K=400)
    ex = B.env_exact(A=A, want_C=False)
    rn = r / t.einsum("nii->n", r)[:, None, None]
    assert (rn - ex.r).abs().max().item() < 1e-9
    assert (ray - 1).abs().max().item() < 1e-9


def test_tm_power_batch_beyond_grid_limit(env):
    """N * d > 65535 (the grid z-limit of the GEMM launches): consecutive chunks on one stream give the
    same result per problem as the small batch."""
    t, B, O = env["torch"], env["B"], env["O"]
    base = t.from_numpy(tensors(4, 5, 91, O)).cuda()
    N = 32768 + 7
    idx = t.arange(N, device="cuda") % 5
    A, Bt = base[idx].contiguous(), base[(idx + 1) % 5].contiguous()
    r, ray = B.tm_power(A, Bt, K=6)
    r5, ray5 = B.tm_power(base, base[(t.arange(5, device="cuda") + 1) % 5].contiguous(), K=6)
    assert (r - r5[idx]).abs().max().item() == 0.0 and (ray - ray5[idx]).abs().max().item() == 0.0
    for dt in (t.complex64,):
        r32, _ = B.tm_power(A.to(dt), Bt.to(dt), K=6)
        assert (r32.to(t.complex128) - r).abs().max().item() < 1e-5


def test_tm_power_complex64(env):
    t, B, O = env["torch"], env["B"], env["O"]
    A, Bt = tensors(64, 2, 464, O), tensors(64, 2, 564, O)
    r, ray = B.tm_power(t.from_numpy(A).cuda().to(t.complex64), t.from_numpy(Bt).cuda().to(t.complex64), K=8)
    for k in range(2):
        r0, q0 = O.power_method(A[k], Bt[k], 8)
        assert np.abs(r[k].cpu().numpy() - r0).max() < 1e-5
        assert abs(ray[k].item() - q0) < 1e-5


# ---------------------------------------------------------------- cfg 5, complex64 on tcgen05
def _cgemm_tc(env, X, Y, conj):
    t, L = env["torch"], env["L"]
    batch, nsum, M, K = X.shape
    N = Y.shape[2]
    Xd, Yd = t.from_numpy(X).cuda(), t.from_numpy(Y).cuda()
    C = t.empty((batch, M, N), dtype=t.complex64, device="cuda")
    L.check(L.load().qmps_cgemm_c64_tc(batch, nsum, M, N, K, Xd.data_ptr(), Yd.data_ptr(), int(conj), C.data_ptr(),
                                       t.cuda.current_stream().cuda_stream), "cgemm_c64_tc")
    t.cuda.synchronize()
    return C.cpu().numpy()


@pytest.mark.parametrize("persistent", [1, 0])
@pytest.mark.parametrize("shape", [(1, 1, 64, 64, 32, 0), (1, 1, 64, 64, 32, 1), (2, 2, 64, 64, 128, 1),
                                   (3, 1, 128, 192, 96, 0), (40, 2, 128, 128, 64, 1), (700, 1, 64, 64, 64, 0)])
def test_cgemm_tc_vs_numpy(env, shape, persistent):
    """The tcgen05 3xTF32 tile kernel against a float64 numpy product: single slab, multi-slab ring
    wrap-around, summands, several tiles per matrix, more tiles than SMs (persistent loop,
    both TMEM accumulators), both op(Y)."""
    batch, nsum, M, N, K, conj = shape
    env["L"].load().qmps_set_option(b"tc_persistent", persistent)
    try:
        rng = np.random.default_rng(batch * 1000 + K + conj)
        X = (rng.normal(size=(batch, nsum, M, K)) + 1j * rng.normal(size=(batch, nsum, M, K))).astype(np.complex64)
        Y = (rng.normal(size=(batch, nsum, N, K)) + 1j * rng.normal(size=(batch, nsum, N, K))).astype(np.complex64)
        ref = np.einsum("btmk,btnk->bmn", X.astype(np.complex128), (Y.conj() if conj else Y).astype(np.complex128))
        got = _cgemm_tc(env, X, Y, conj)
        assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5          # complex64 tolerance of the north star
    finally:
        env["L"].load().qmps_set_option(b"tc_persistent", 1)


def test_cgemm_tc_linearity_and_edge_cases(env):
    """Size-independent properties at the cfg-5 size (D = 256): C(X1 + X2, Y) = C(X1, Y) + C(X2, Y),
    C(X, Y)^H = C(Y, X) for op = conjugate transpose; empty batch; unsupported shapes are refused."""
    t, L = env["torch"], env["L"]
    rng = np.random.default_rng(5)
    mk = lambda *s: (rng.normal(size=s) + 1j * rng.normal(size=s)).astype(np.complex64)
    X1, X2, Y = mk(4, 2, 256, 256), mk(4, 2, 256, 256), mk(4, 2, 256, 256)
    c1, c2, c12 = _cgemm_tc(env, X1, Y, 1), _cgemm_tc(env, X2, Y, 1), _cgemm_tc(env, X1 + X2, Y, 1)
    scale = np.abs(c12).max()
    assert np.abs(c12 - (c1 + c2)).max() / scale < 1e-5
    assert np.abs(_cgemm_tc(env, Y, X1, 1) - np.conj(np.swapaxes(c1, -1, -2))).max() / scale < 1e-5
    lib = L.load()
    assert lib.qmps_cgemm_c64_tc(0, 1, 64, 64, 32, None, None, 0, None, None) == 0
    z = t.zeros((1, 1, 96, 32), dtype=t.complex64, device="cuda")
    assert lib.qmps_cgemm_c64_tc(1, 1, 96, 64, 32, z.data_ptr(), z.data_ptr(), 0, z.data_ptr(), None) != 0   # M % 64


@pytest.mark.parametrize("D,cnt,K", [(64, 3, 8), (128, 2, 5), (256, 2, 8), (64, 2, 0), (64, 2, 1)])
def test_tm_power_complex64_tcgen05_vs_oracle(env, D, cnt, K):
    """cfg 5 in complex64 runs on tcgen05 for D % 64 == 0: same K steps as the oracle, 1e-5."""
    t, B, O = env["torch"], env["B"], env["O"]
    A, Bt = tensors(D, cnt, 1400 + D, O), tensors(D, cnt, 1500 + D, O)
    rng = np.random.default_rng(D + K)
    r0 = rng.normal(size=(cnt, D, D)) + 1j * rng.normal(size=(cnt, D, D))
    r0 /= np.linalg.norm(r0, axis=(1, 2), keepdims=True)
    c64 = lambda x: t.from_numpy(x).cuda().to(t.complex64)
    for start in (None, r0):
        r, ray = B.tm_power(c64(A), c64(Bt), K=K, r0=None if start is None else c64(start))
        for k in range(cnt):
            ref_r, ref_q = O.power_method(A[k], Bt[k], K, None if start is None else start[k])
            assert np.abs(r[k].cpu().numpy() - ref_r).max() < 1e-5
            assert abs(ray[k].item() - ref_q) < 1e-5


def test_tm_power_complex64_tcgen05_matches_simt_path(env):
    """The tensor-core path and the SIMT complex64 path agree (same inputs, option toggled)."""
    t, B, O, L = env["torch"], env["B"], env["O"], env["L"]
    A, Bt = tensors(64, 4, 1464, O), tensors(64, 4, 1564, O)
    c64 = lambda x: t.from_numpy(x).cuda().to(t.complex64)
    r1, q1 = B.tm_power(c64(A), c64(Bt), K=16)
    L.load().qmps_set_option(b"tc_power", 0)
    try:
        r2, q2 = B.tm_power(c64(A), c64(Bt), K=16)
    finally:
        L.load().qmps_set_option(b"tc_power", 1)
    assert (r1 - r2).abs().max().item() < 1e-5 and (q1 - q2).abs().max().item() < 1e-5


def test_argmin(env):
    t, B = env["torch"], env["B"]
    rng = np.random.default_rng(0)
    for n in (1, 257, 100003):
        c = rng.normal(size=n)
        c[n // 3] = c.min() - 1
        c[(2 * n) // 3] = c.min()                          # tie: the first index wins
        bc, bi = B.argmin(t.from_numpy(c).cuda(), 1000)
        assert bc.item() == c.min() and bi.item() == 1000 + int(np.argmin(c))


# ---------------------------------------------------------------- drop-in numpy API
def test_dropin_tools_against_reference_golden(env, golden):
    from qmps_b200 import tools, time_evolve_tools as tet
    O = env["O"]
    g = golden["ref_tools"]
    for D in (2, 4):
        for U, Vref in zip(g[f"u2t_U_D{D}"], g[f"env_V_D{D}"]):
            V = tools.get_env_exact(U)
            assert np.abs(V[:, 0] - Vref[:, 0]).max() < 1e-10       # the unique column of the reference's V
            assert np.abs(V.conj().T @ V - np.eye(D * D)).max() < 1e-12
            assert np.array_equal(tools.unitary_to_tensor(U), O.unitary_to_tensor(U))
    U, passed = tools.tensor_to_unitary(g["u2t_A_D2"][0], testing=True)
    assert passed
    ext = tools.unitary_extension(g["uext_tall_in"])
    assert np.abs(ext[:, :3] - g["uext_tall_in"]).max() == 0 and np.abs(ext.conj().T @ ext - np.eye(6)).max() < 1e-13
    ext = tools.unitary_extension(g["uext_wide_in"])
    assert np.abs(ext[:3] - g["uext_wide_in"]).max() == 0 and np.abs(ext @ ext.conj().T - np.eye(6)).max() < 1e-13
    assert tools.unitary_extension(g["uext_tall_in"], 8).shape == (8, 8)
    assert np.array_equal(tools.environment_from_unitary(g["e2u_out"]), g["efu_out"])
    # environments on a site: round trips (reference self-tests, new_time_evolve.py:53-70)
    rng = np.random.default_rng(1)
    q = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    for put, off in ((tet.put_env_on_left_site, tet.get_env_off_left_site), (tet.put_env_on_right_site, tet.get_env_off_right_site)):
        Aq, n = put(q, ret_n=True)
        assert np.abs(Aq.conj().T @ Aq - np.eye(4)).max() < 1e-13
        assert np.abs(off(Aq) * n - q).max() < 1e-13
    p1, p2 = rng.normal(size=15), rng.normal(size=15)
    f, r = tet.get_overlap_exact(p1, p2)
    f0, r0 = O.get_overlap_exact(p1, p2)
    assert abs(f - f0) < 1e-10
    assert abs(tet.get_overlap_exact(p1, p1, testing=False) - 1) < 1e-12
    A, Bt = tensors(2, 2, 77, O)
    assert np.abs(tet.merge(A, Bt) - O.merge(A, Bt)).max() < 1e-14


def test_dropin_ground_state(env):
    from qmps_b200 import ground_state as gs
    O = env["O"]
    H = gs.Hamiltonian({"ZZ": -1, "X": 1}).to_matrix()
    assert np.array_equal(H, O.tfim_matrix(1.0))
    opt = gs.SparseFullEnergyOptimizer(H, D=2, depth=2, initial_guess=np.array([0.3, -0.2, 0.5, 0.1]))
    e = opt.objective_function(opt.initial_guess)
    assert abs(e - O.energy_transfer(O.unitary_to_tensor(O.shallow_cnot_state_tensor(2, opt.initial_guess)), H)) < 1e-10
    rng = np.random.default_rng(3)
    p = rng.normal(size=15)
    opt = gs.NonSparseFullEnergyOptimizer(H, D=2, initial_guess=p)
    assert abs(opt.objective_function(p) - O.energy_transfer(O.unitary_to_tensor(gs.SU(p, 4)), H)) < 1e-10
    p = rng.normal(size=30)
    opt2 = gs.NonSparseFullTwoSiteEnergyOptimizer(H, initial_guess=p)
    ref = O.energy_two_site_statevector(gs.SU(p[:15], 4), gs.SU(p[15:], 4), H)
    assert abs(opt2.objective_function(p) - ref) < 1e-10


def test_env_exact_host_entry_point(env):
    """The host-buffer C ABI call (what bench.py's e2e leg times)."""
    import ctypes
    L, O = env["L"], env["O"]
    N = 5000
    A = np.tile(tensors(2, 50, 600, O), (100, 1, 1, 1))
    eta = np.empty(N, np.complex128); r = np.empty((N, 2, 2), np.complex128)
    C = np.empty((N, 2, 2), np.complex128); st = np.empty(N, np.int32)
    rc = L.load().qmps_env_exact_host(2, 2, N, A.ctypes.data, 0, 1, eta.ctypes.data, r.ctypes.data, C.ctypes.data,
                                      st.ctypes.data, L.C128, 0)
    L.check(rc, "env_exact_host")
    assert st.sum() == 0
    for k in range(50):
        _, r0, C0, _ = O.env_exact_parts(A[k])
        assert np.abs(r[k] - r0).max() < 1e-10 and np.abs(r[k + 50 * 99] - r0).max() < 1e-10
        assert np.abs(C[k] - C0).max() < 1e-9


# ---------------------------------------------------------------- canonical forms, Es, overlap
_PAULIS = np.array([[[0, 1], [1, 0]], [[0, -1j], [1j, 0]], [[1, 0], [0, -1]]], dtype=complex)


def _random_tensors(d, D, count, seed):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.normal(size=(count, d, D, D)) + 1j * rng.normal(size=(count, d, D, D)))


@pytest.mark.parametrize("d,D,count", [(2, 2, 133), (2, 4, 40), (4, 4, 9), (2, 8, 6), (2, 16, 2)])
def test_left_canonicalise_and_mixed_vs_oracle(env, d, D, count):
    """iMPS([A]).left_canonicalise() / .mixed() on random NON-canonical tensors (ragged batch sizes)
    against oracle/canonical.py, plus the properties of the reference's tests/test_represent.py:23-31."""
    t, B, O = env["torch"], env["B"], env["O"]
    A = _random_tensors(d, D, count, 900 + D + d)
    lc = B.left_canonicalise(t.from_numpy(A).cuda(), want_L=True)
    mx = B.mixed_canonical(t.from_numpy(A).cuda())
    assert int(lc.status.abs().sum()) == 0 and int(mx.status.abs().sum()) == 0
    AL, Lm, eta = lc.AL.cpu().numpy(), lc.L.cpu().numpy(), lc.eta.cpu().numpy()
    AL2, AR, C = mx.AL.cpu().numpy(), mx.AR.cpu().numpy(), mx.C.cpu().numpy()
    assert np.array_equal(AL, AL2)
    I = np.eye(D)
    for k in range(count):
        AL0, AR0, C0 = O.mixed(A[k])
        eta0 = O.eigs(A[k])[0]
        w = np.sort(np.abs(np.linalg.eigvals(O.transfer_matrix(A[k]))))[::-1]
        tol = 10 * TOL / min(1.0, 1 - w[1] / w[0]) * np.linalg.cond(Lm[k])
        assert abs(eta[k] - eta0) < TOL * abs(eta0)
        assert np.abs(AL[k] - AL0).max() < tol
        assert np.abs(C[k] - C0).max() < 10 * tol and np.abs(AR[k] - AR0).max() < 100 * tol
        assert np.abs(np.einsum("sij,sik->jk", AL[k].conj(), AL[k]) - I).max() < 1e-11     # left-canonical
        assert np.abs(np.einsum("sij,skj->ik", AR[k], AR[k].conj()) - I).max() < 1e-8     # right-canonical
        assert np.abs(np.tril(Lm[k], -1)).max() == 0 and abs(np.trace(Lm[k].conj().T @ Lm[k]) - D) < 1e-10
        rr, ll = C[k] @ C[k].conj().T, C[k].conj().T @ C[k]
        EL, ER = O.transfer_matrix(AL[k]), O.transfer_matrix(AR[k])
        assert np.abs(EL @ rr.reshape(-1) - rr.reshape(-1)).max() < 1e-10
        assert np.abs(ER.conj().T @ ll.reshape(-1) - ll.reshape(-1)).max() < 1e-8


def test_mixed_of_left_canonical_input_is_env_exact(env):
    """assume_left_canonical: AL = A, C is exactly the Cholesky factor get_env_exact uses
    (qmps/tools.py:184-186 get_env_exact_alternative == get_env_exact up to gauge)."""
    t, B, O = env["torch"], env["B"], env["O"]
    for D, count in ((2, 257), (4, 31)):
        A = t.from_numpy(tensors(D, count, 950, O)).cuda()
        mx = B.mixed_canonical(A, assume_left_canonical=True)
        ex = B.env_exact(A=A)
        assert int(mx.status.abs().sum()) == 0
        assert t.equal(mx.C, ex.C) and mx.AL.data_ptr() == A.data_ptr()
        AR = mx.AR.cpu().numpy()
        for k in range(0, count, 7):
            assert O.is_right_canonical(AR[k], 1e-9)


@pytest.mark.parametrize("D,count", [(2, 100), (4, 20), (8, 5)])
def test_expectation_values_vs_oracle(env, D, count):
    """iMPS.Es: gauge invariant, so the canonicalised tensor, the original tensor (general formula)
    and the oracle must agree; for unitary-derived tensors also the energy identity
    <O x 1> = Es(O) (qmps/ground_state.py:251-266 evaluated with H = O (x) 1)."""
    t, B, O = env["torch"], env["B"], env["O"]
    A = _random_tensors(2, D, count, 970 + D)
    Ad = t.from_numpy(A).cuda()
    es_gen = B.expectation_values(Ad, _PAULIS, assume_left_canonical=False).cpu().numpy()
    AL = B.left_canonicalise(Ad).AL
    es_can = B.expectation_values(AL, _PAULIS).cpu().numpy()
    for k in range(count):
        ref = O.expectation_values(A[k], _PAULIS)
        w = np.sort(np.abs(np.linalg.eigvals(O.transfer_matrix(A[k]))))[::-1]
        tol = 100 * TOL / min(1.0, 1 - w[1] / w[0])
        assert np.abs(es_gen[k] - ref).max() < tol and np.abs(es_can[k] - ref).max() < tol
        assert np.abs(es_can[k].imag).max() < 1e-12
    Au = t.from_numpy(tensors(D, count, 980, O)).cuda()
    es = B.expectation_values(Au, _PAULIS).cpu().numpy()
    for o in range(3):
        H = np.kron(_PAULIS[o], np.eye(2))
        e = B.energy_tensor(Au, H).cpu().numpy()
        assert np.abs(es[:, o].real - e).max() < 1e-11


def test_expectation_complex64_and_empty(env):
    t, B, O = env["torch"], env["B"], env["O"]
    A = tensors(4, 50, 990, O)
    e64 = B.expectation_values(t.from_numpy(A).cuda().to(t.complex64), _PAULIS).cpu().numpy()
    e128 = B.expectation_values(t.from_numpy(A).cuda(), _PAULIS).cpu().numpy()
    assert np.abs(e64 - e128).max() < 1e-4
    lc = B.left_canonicalise(t.from_numpy(_random_tensors(2, 4, 20, 991)).cuda().to(t.complex64))
    AL = lc.AL.cpu().numpy().astype(complex)
    assert np.abs(np.einsum("nsij,nsik->njk", AL.conj(), AL) - np.eye(4)).max() < 1e-4
    empty = t.empty((0, 2, 4, 4), dtype=t.complex128, device="cuda")
    assert B.left_canonicalise(empty).AL.shape == (0, 2, 4, 4)
    assert B.mixed_canonical(empty).C.shape == (0, 4, 4)


def test_imps_compat_loop_quantities(env):
    """The three per-step quantities of the reference's Loschmidt loop
    (qmps/loschmidts/time_evo.py:143-145): left_canonicalise, Es(paulis), overlap."""
    t, O = env["torch"], env["O"]
    from qmps_b200.imps import iMPS, Map, TransferMatrix
    rng = np.random.default_rng(5)
    A0 = O.unitary_to_tensor(O.shallow_full_state_tensor(rng.normal(size=15)))
    A1 = O.unitary_to_tensor(O.shallow_full_state_tensor(rng.normal(size=15)))
    A_ = iMPS([A1]).left_canonicalise()
    assert O.is_left_canonical(A_[0])
    assert np.abs(A_.Es(_PAULIS) - O.expectation_values(A1, _PAULIS).real).max() < 1e-10
    assert abs(A_.overlap(iMPS([A0])) - O.overlap(A1, A0)) < 1e-10
    assert abs(A_.overlap(A_) - 1) < 1e-10
    AL, AR, C = iMPS([rng.normal(size=(2, 2, 2)) + 1j * rng.normal(size=(2, 2, 2))]).mixed()
    r, l, I = C @ C.conj().T, C.conj().T @ C, np.eye(2)
    assert Map(AL[0], AL[0]).is_right_eigenvector(r) and Map(AL[0], AL[0]).is_left_eigenvector(I)
    assert Map(AR[0], AR[0]).is_right_eigenvector(I) and Map(AR[0], AR[0]).is_left_eigenvector(l)
    eta, l2, r2 = TransferMatrix(AL[0]).eigs()
    assert abs(eta - 1) < 1e-10 and np.abs(r2 - r).max() < 1e-9 and abs(np.trace(l2 @ r2) - 1) < 1e-10


# ---------------------------------------------------------------- BASELINE configs at FULL size: properties
def _tfim_gates(NT, dt=0.02, g=0.2):
    from scipy.linalg import expm
    from qmps_b200.ground_state import Hamiltonian
    H = Hamiltonian({'ZZ': -1, 'X': g}).to_matrix()
    return np.stack([expm(-1j * H * 2 * dt * k) for k in range(NT)])


@pytest.mark.parametrize("D,P,gate", [(4, 12, "cnot"), (2, 15, "full")])
def test_loschmidt_full_size_properties(env, D, P, gate):
    """BASELINE config 3 (D = 4) and its D = 2 twin at full size, 4096 parameter sets x 1000 times:
    |eta| <= 1 everywhere, at t = 0 (W = 1) the cost of parameter set 0 against itself is exactly -1
    and every other one is -sqrt(per-site fidelity) of the two states; sampled entries against the
    oracle; statuses clean."""
    t, B, O, R = env["torch"], env["B"], env["O"], env["R"]
    NP, NT = 4096, 1000
    theta = np.random.default_rng(2).normal(size=(NP, P))
    prog = (R.ShallowCNOTStateTensor_nonuniform(4, np.zeros(P)) if gate == "cnot" else R.ShallowFullStateTensor(2, np.zeros(P))).program()
    th = t.from_numpy(theta).cuda()
    A0 = B.ansatz_tensors(prog, th[:1])[0]
    W = _tfim_gates(NT)
    cost, echo, eta = B.loschmidt_costs(prog, th, A0, t.from_numpy(W).cuda())
    t.cuda.synchronize()
    assert cost.shape == (NP, NT)
    assert float(eta.abs().max()) <= 1 + 1e-10
    assert abs(float(cost[0, 0]) + 1) < 1e-10
    fid = B.overlap_theta(prog, th[:1].expand(NP, P).contiguous(), th).fid          # |eta(E_{A0 B_p})|^2
    assert float((cost[:, 0] + fid.sqrt()).abs().max()) < 1e-9                     # two-site map at W = 1: eta_2 = eta^2, |eta_2| = fid
    assert float((echo + 2 * t.log(-cost) * 2).abs().max()) < 1e-8                 # echo = -log|eta|^2, cost = -sqrt|eta|
    A0n = A0.cpu().numpy()
    rs = np.random.RandomState(0)
    for _ in range(12):
        p, k = int(rs.randint(NP)), int(rs.randint(NT))
        Bp = B.ansatz_tensors(prog, th[p:p + 1])[0].cpu().numpy()
        assert abs(float(cost[p, k]) - O.loschmidt_cost(A0n, Bp, W[k])) < TOL * 10


def test_rotosolve_full_size_properties(env):
    """BASELINE config 4 at full size (65536 parameter vectors, D = 8, 3 shifts on one coordinate): the
    shift-0 column equals the plain energy, every energy lies inside the spectrum of the Heisenberg bond,
    the closed-form coordinate update equals the reference formula on every vector and touches no other
    coordinate, and the device argmin of the re-evaluated energies equals torch's."""
    t, B, R = env["torch"], env["B"], env["R"]
    from qmps_b200.ground_state import Hamiltonian
    N, P, coord = 65536, 24, 5
    theta = t.from_numpy(np.random.default_rng(3).normal(size=(N, P))).cuda()
    prog = R.ShallowCNOTStateTensor_nonuniform(8, np.zeros(P)).program()
    H = Hamiltonian({'XX': 1, 'YY': 1, 'ZZ': 1}).to_matrix()
    e3, st = B.energy_theta(prog, theta, H, coord=coord, shifts=B.ROTO3_SHIFTS, want_status=True)
    e0 = B.energy_theta(prog, theta, H)
    t.cuda.synchronize()
    assert int(st.abs().sum()) == 0
    assert float((e3[:, 0] - e0).abs().max()) < 1e-11
    w = np.linalg.eigvalsh(H)
    assert float(e3.min()) >= w[0] - 1e-9 and float(e3.max()) <= w[-1] + 1e-9
    # the closed-form coordinate update of qmps/rotosolve.py:175-177 on all 65536 vectors, against numpy.  (For an
    # iMPS the energy is NOT a single sinusoid in one parameter -- the same gate sits on every site and in the
    # environment -- so "theta* is the minimum" is not a property of the path; the reference applies the formula anyway.)
    new = theta.clone()
    B.rotosolve_fit(e3, new, coord)
    e_new = B.energy_theta(prog, new, H)
    e = e3.cpu().numpy()
    wrap = lambda x: np.arctan2(np.sin(x), np.cos(x))                     # noqa: E731
    step = wrap(-np.pi / 2 - np.arctan2(2 * e[:, 0] - e[:, 1] - e[:, 2], e[:, 1] - e[:, 2]))
    want = wrap(theta[:, coord].cpu().numpy() + step)
    assert np.abs(new[:, coord].cpu().numpy() - want).max() < 1e-12
    changed = (new - theta).abs().sum(dim=1)
    assert float((new - theta)[:, [c for c in range(P) if c != coord]].abs().max()) == 0 and float(changed.max()) > 0
    bc, bi = B.argmin(e_new)
    assert float(bc) == float(e_new.min()) and int(bi) == int(t.argmin(e_new))


@pytest.mark.parametrize("dt_name", ["complex128", "complex64"])
def test_tm_power_full_size_properties(env, dt_name):
    """BASELINE config 5 at full size (D = 64 x 512 and D = 256 x 32 problems, K = 32): for A = B
    left-canonical the Rayleigh quotient converges to 1, r stays unit-Frobenius and
    Hermitian, and applying the map once more by an independent einsum reproduces eta r."""
    t, B = env["torch"], env["B"]
    cdt = getattr(t, dt_name)
    tol = 1e-9 if cdt == t.complex128 else 2e-4
    for D, N in ((64, 512), (256, 32)):
        g = t.Generator(device="cuda").manual_seed(40 + D)
        Z = t.randn((N, 2 * D, D), dtype=t.float64, device="cuda", generator=g) + 1j * t.randn((N, 2 * D, D), dtype=t.float64, device="cuda", generator=g)
        Q, _ = t.linalg.qr(Z)
        A = Q.reshape(N, D, 2, D).permute(0, 2, 1, 3).contiguous().to(cdt)
        r, ray = B.tm_power(A, A, K=32)
        t.cuda.synchronize()
        assert float((t.linalg.matrix_norm(r) - 1).abs().max()) < tol * 10
        assert float((r - r.conj().transpose(1, 2)).abs().max()) < tol * 10
        assert float((ray - 1).abs().max()) < 1e-4                       # 32 applications of a gapped CPTP map
        r64, A64 = r[:4].to(t.complex128), A[:4].to(t.complex128)
        Er = t.einsum("nsij,njl,nskl->nik", A64, r64, A64.conj())
        q = t.einsum("nij,nij->n", r64.conj(), Er)
        assert float((q - ray[:4].to(t.complex128)).abs().max()) < tol * 10


def test_energy_theta_against_reference_ground_state_script(env, golden):
    """scripts/ground_state_finding.py:83-128 (the reference's cirq-free energy route) as a gate program:
    Rx Rx / Rz Rz / CNOT layers.  The CUDA ansatz evaluator reproduces the script's unitaries (qubit order,
    rotation conventions) and the fused energy kernel its energies, for 1, 2 and 4 layers."""
    t, B, R, O = env["torch"], env["B"], env["R"], env["O"]
    g = golden["ref_ground_state_script"]
    for layers in (1, 2, 4):
        P = 4 * layers
        prog = R.GateProgram(2, P)
        for l in range(layers):
            prog.rx(0, 4 * l).rx(1, 4 * l + 1).rz(0, 4 * l + 2).rz(1, 4 * l + 3).cnot(0, 1)
        theta = g[f"p_L{layers}"]
        U = B.ansatz_unitaries_host(prog, theta)
        assert np.abs(U[:, :, :2] - g[f"U_L{layers}"][:, :, :2]).max() < 1e-13      # the columns the tensor uses
        assert np.abs(U - g[f"U_L{layers}"]).max() < 1e-13
        for lam in (0.5, 1.0):
            e = B.energy_theta(prog, theta, O.tfim_matrix(lam)).cpu().numpy()
            assert np.abs(e - g[f"eps_L{layers}_lam{lam}"]).max() < 1e-10
            e32 = B.energy_theta(prog, theta, O.tfim_matrix(lam), dtype=t.complex64).cpu().numpy()
            assert np.abs(e32 - g[f"eps_L{layers}_lam{lam}"]).max() < 2e-5


def test_dropin_time_evolve_tools_against_reference_functions(env, golden):
    """qmps_b200.time_evolve_tools (merge and the environment embeddings, which complete their unitaries on
    the GPU) against the reference's own functions cut out of qmps/time_evolve_tools.py:20-74
    (tests/golden/ref_misc.npz)."""
    from qmps_b200 import time_evolve_tools as TE
    g = golden["ref_misc"]
    for a, b, m in zip(g["merge_A"], g["merge_B"], g["merge_out"]):
        assert np.abs(TE.merge(a, b) - m).max() < 1e-14
    for k, q in enumerate(g["env_q"]):
        UL, nL = TE.put_env_on_left_site(q, ret_n=True)
        UR, nR = TE.put_env_on_right_site(q, ret_n=True)
        assert abs(nL - g["left_n"][k]) < 1e-13 and abs(nR - g["right_n"][k]) < 1e-13
        assert np.abs(TE.get_env_off_left_site(UL) - g["left_off"][k]).max() < 1e-12
        assert np.abs(TE.get_env_off_right_site(UR) - g["right_off"][k]).max() < 1e-12
        assert np.abs(UR[:2] - g["right_U"][k][:2]).max() < 1e-12                 # the defining rows are unique
        for U in (UL, UR):
            assert np.abs(U @ U.conj().T - np.eye(4)).max() < 1e-12


def test_dropin_loschmidt_obj_against_reference_function(env, golden):
    """qmps_b200.loschmidts.time_evo.obj / obj_batched (ansatz -> merge -> D = 2 register eigen-solver, all on
    the GPU) against the reference's own `obj(p, A, WW)` (qmps/loschmidts/time_evo.py:75-116) executed
    unmodified by oracle/make_golden_obj.py."""
    t = env["torch"]
    from qmps_b200.loschmidts import time_evo as TE
    g = golden["ref_loschmidt_obj"]
    for a in range(len(g["A0"])):
        cost, echo = TE.obj_batched(g["ps"], g["A0"][a], g["Ws"])
        assert np.abs(cost.cpu().numpy() - g["obj"][a]).max() < 1e-10
        assert np.abs(echo.cpu().numpy() + 4 * np.log(-g["obj"][a])).max() < 1e-8
    assert abs(TE.obj(g["ps"][1], g["A0"][0], g["Ws"][1]) - g["obj"][0, 1, 1]) < 1e-10
    # complex64 mode through the batched pipeline
    from qmps_b200 import batched as B, represent as R
    prog = R.ShallowFullStateTensor(2, np.zeros(15)).program()
    c32 = B.loschmidt_costs(prog, t.from_numpy(g["ps"]).cuda(), t.from_numpy(g["A0"][0]).cuda().to(t.complex64),
                            t.from_numpy(g["Ws"]).cuda().to(t.complex64), dtype=t.complex64)[0]
    assert np.abs(c32.cpu().numpy() - g["obj"][0]).max() < 1e-5


def test_dropin_get_overlap_exact_against_reference_function(env, golden):
    """qmps_b200.time_evolve_tools.get_overlap_exact (D = 2 register eigen-solver with eigenvector) against the
    reference's own function (qmps/time_evolve_tools.py:84-91, oracle/make_golden_obj.py): the fidelity to 1e-10;
    r is a right eigenvector of the mixed map (the reference's r is in xmps' gauge, supplied by the oracle)."""
    from qmps_b200 import time_evolve_tools as TE
    O = env["O"]
    g = golden["ref_loschmidt_obj"]
    for a in range(3):
        for b in range(4):
            f, r = TE.get_overlap_exact(g["p0"][a], g["ps"][b])
            assert abs(f - g["overlap"][a, b]) < 1e-10
            A = O.unitary_to_tensor(O.shallow_full_state_tensor(g["p0"][a]))
            Bt = O.unitary_to_tensor(O.shallow_full_state_tensor(g["ps"][b]))
            E = O.transfer_matrix(A, Bt)
            v = np.asarray(r).reshape(-1)
            lam = (v.conj() @ (E @ v)) / (v.conj() @ v)
            assert abs(abs(lam) ** 2 - f) < 1e-9 and np.abs(E @ v - lam * v).max() < 1e-9


def test_torch_library_ops_match_batched_api(env):
    """The torch.library layer (qmps_b200/ops.py) returns exactly what the batched API returns."""
    t, B, O = env["torch"], env["B"], env["O"]
    import qmps_b200.ops  # noqa: F401
    from qmps_b200 import brickwall as BW
    A = t.from_numpy(tensors(4, 9, 61, O)).cuda()
    Bt = t.from_numpy(tensors(4, 5, 62, O)).cuda()
    eta, r, C, st = t.ops.qmps_b200.env_exact(A)
    ref = B.env_exact(A=A)
    assert t.equal(r, ref.r) and t.equal(C, ref.C) and t.equal(eta, ref.eta) and t.equal(st, ref.status)
    e2, cost, echo, fid = t.ops.qmps_b200.fixed_point_cost(A, Bt, True)
    fp = B.fixed_point(A, Bt, pair="outer", want_vec=False)
    assert cost.shape == (9, 5) and t.equal(cost, fp.cost) and t.equal(fid, fp.fid)
    rK, ray = t.ops.qmps_b200.tm_power(A, A, 3)
    r2, ray2 = B.tm_power(A, A, 3)
    assert t.equal(rK, r2) and t.equal(ray, ray2)
    U = t.from_numpy(haar_batch(4, 6, 71)).cuda()
    W = t.from_numpy(haar_batch(16, 1, 72)[0]).cuda()
    c = t.ops.qmps_b200.bw_evolve_cost(U[0], U[1], U[2:], U[2:].flip(0).contiguous(), W)
    assert t.equal(c, BW.bw_evolve_cost(U[0], U[1], U[2:], U[2:].flip(0).contiguous(), W))


# ---------------------------------------------------------------- complex128 on tcgen05 kind::i8
@pytest.mark.parametrize("shape", [(1, 64, 32, 64, 0), (2, 64, 64, 128, 1), (3, 128, 96, 64, 0), (2, 256, 256, 256, 1), (1, 64, 32, 512, 1)])
def test_zgemm_i8_vs_numpy(env, shape):
    """The int8-slice complex128 product (kernels_tc_i8.cuh) against numpy: random matrices, rows of very
    different magnitude (the per-row power-of-two scales), exact zeros, and a structured case."""
    t, B = env["torch"], env["B"]
    batch, M, N, K, conj = shape
    rng = np.random.default_rng(M + N + K)
    X = rng.normal(size=(batch, M, K)) + 1j * rng.normal(size=(batch, M, K))
    Y = rng.normal(size=(batch, N, K)) + 1j * rng.normal(size=(batch, N, K))
    X *= np.exp2(rng.integers(-30, 30, size=(batch, M, 1)))          # row scales over 18 orders of magnitude
    Y *= np.exp2(rng.integers(-30, 30, size=(batch, N, 1)))
    X[0, 3] = 0.0; Y[0, 5].imag = 0.0
    C = B.zgemm_i8(X, Y, conj_y=bool(conj)).cpu().numpy()
    ref = X @ (Y.conj() if conj else Y).transpose(0, 2, 1)
    # error bound of the scheme: 2^-42 of |row|max |col|max K per entry; compare entrywise against that scale
    scale = np.maximum(np.abs(X).max(axis=2)[:, :, None] * np.abs(Y).max(axis=2)[:, None, :] * K, 1e-300)
    assert (np.abs(C - ref) / scale).max() < 1e-11
    assert np.abs(C[0, 3]).max() == 0.0
    # relative Frobenius error on well-scaled operands: the 4e-12 of the numpy prototype
    X1 = rng.normal(size=(batch, M, K)) + 1j * rng.normal(size=(batch, M, K))
    Y1 = rng.normal(size=(batch, N, K)) + 1j * rng.normal(size=(batch, N, K))
    C1 = B.zgemm_i8(X1, Y1, conj_y=bool(conj)).cpu().numpy()
    ref1 = X1 @ (Y1.conj() if conj else Y1).transpose(0, 2, 1)
    assert np.linalg.norm(C1 - ref1) / np.linalg.norm(ref1) < 2e-11
    # integers are reproduced exactly
    Xi = rng.integers(-50, 50, size=(batch, M, K)) + 1j * rng.integers(-50, 50, size=(batch, M, K))
    Yi = rng.integers(-50, 50, size=(batch, N, K)) + 1j * rng.integers(-50, 50, size=(batch, N, K))
    Ci = B.zgemm_i8(Xi, Yi, conj_y=bool(conj)).cpu().numpy()
    assert np.array_equal(Ci, Xi @ (Yi.conj() if conj else Yi).transpose(0, 2, 1))


@pytest.mark.parametrize("D,cnt,K", [(64, 3, 8), (128, 2, 5), (256, 2, 8), (64, 2, 0), (64, 2, 1)])
def test_tm_power_complex128_tcgen05_vs_oracle_and_dmma(env, D, cnt, K):
    """complex128 power method on tcgen05 kind::i8 against the oracle doing the identical K steps (1e-10), and
    against the FP64 tensor-pipe (DMMA) path it replaces."""
    t, B, O, L = env["torch"], env["B"], env["O"], env["L"]
    A, Bt = tensors(D, cnt, 4000 + D, O), tensors(D, cnt, 5000 + D, O)
    lib = L.load()
    lib.qmps_set_option(b"i8_power", 2)          # force the int8 path (the default picks it from D = 128 up)
    r, ray = B.tm_power(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), K)
    lib.qmps_set_option(b"i8_power", 0)
    try:
        r2, ray2 = B.tm_power(t.from_numpy(A).cuda(), t.from_numpy(Bt).cuda(), K)
        t.cuda.synchronize()
    finally:
        lib.qmps_set_option(b"i8_power", 1)
    r, ray, r2, ray2 = r.cpu().numpy(), ray.cpu().numpy(), r2.cpu().numpy(), ray2.cpu().numpy()
    for k in range(cnt):
        r0, q0 = O.power_method(A[k], Bt[k], K)
        assert np.abs(r[k] - r0).max() < TOL and abs(ray[k] - q0) < TOL
    assert np.abs(r - r2).max() < TOL and np.abs(ray - ray2).max() < TOL
