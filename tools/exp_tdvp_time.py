import sys, time, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
from qmps_b200 import batched as B, _lib as L
L.require_device()
dev = torch.device('cuda', 0)
D, N = 64, 64
g = torch.Generator(device=dev).manual_seed(64)
Z = torch.randn((N, 2 * D, D), dtype=torch.float64, device=dev, generator=g) + 1j * torch.randn((N, 2 * D, D), dtype=torch.float64, device=dev, generator=g)
Q, _ = torch.linalg.qr(Z)
A = Q.reshape(N, D, 2, D).permute(0, 2, 1, 3).contiguous()
h = torch.tensor(np.kron(np.diag([1.0, -1.0]), np.diag([1.0, -1.0])) * -1.0 + 0.35 * (np.kron([[0, 1], [1, 0]], np.eye(2)) + np.kron(np.eye(2), [[0, 1], [1, 0]])), dtype=torch.complex128, device=dev)
for rep in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = B.tdvp_tangent_large(A, h)
    torch.cuda.synchronize(); print(rep, round((time.perf_counter() - t0) * 1e3, 2), 'ms', out[2]['r_iterations'], out[2]['k_iterations'], flush=True)
