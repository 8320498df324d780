"""``torch.library`` custom ops over the C ABI (the "thin C-ABI / PyTorch custom-op layer" of the north star).

The ops take and return CUDA tensors, enqueue on the current stream and never synchronise; they are registered for
the CUDA dispatch key only -- calling them with CPU tensors raises (there is no CPU implementation):

    torch.ops.qmps_b200.env_exact(A)                         -> (eta[N], r[N,D,D], C[N,D,D], status[N])
    torch.ops.qmps_b200.fixed_point_cost(A, B, outer)        -> (eta, cost, echo, fid)      # Map(A,B) leading eigenvalue
    torch.ops.qmps_b200.tm_power(A, B, K)                    -> (r_K[N,D,D], rayleigh[N])
    torch.ops.qmps_b200.bw_evolve_cost(U1, U2, V1, V2, W)    -> cost[N]

``register_fake`` shape functions are provided, so the ops trace under ``torch.compile`` / ``make_fx`` as opaque
nodes (the hot path itself is never compiled by torch)."""
import torch

from . import batched, brickwall

_RD = {torch.complex128: torch.float64, torch.complex64: torch.float32}

lib = torch.library.Library("qmps_b200", "DEF")
lib.define("env_exact(Tensor A) -> (Tensor, Tensor, Tensor, Tensor)")
lib.define("fixed_point_cost(Tensor A, Tensor B, bool outer) -> (Tensor, Tensor, Tensor, Tensor)")
lib.define("tm_power(Tensor A, Tensor B, int K) -> (Tensor, Tensor)")
lib.define("bw_evolve_cost(Tensor U1, Tensor U2, Tensor V1, Tensor V2, Tensor W) -> Tensor")


def _env_exact(A):
    r = batched.env_exact(A=A)
    return r.eta, r.r, r.C, r.status


def _fixed_point_cost(A, B, outer):
    fp = batched.fixed_point(A, B, pair="outer" if outer else "elementwise", want_vec=False, want_status=False)
    return fp.eta, fp.cost, fp.echo, fp.fid


def _tm_power(A, B, K):
    return batched.tm_power(A, B, K)


def _bw_evolve_cost(U1, U2, V1, V2, W):
    return brickwall.bw_evolve_cost(U1, U2, V1, V2, W)


lib.impl("env_exact", _env_exact, "CUDA")
lib.impl("fixed_point_cost", _fixed_point_cost, "CUDA")
lib.impl("tm_power", _tm_power, "CUDA")
lib.impl("bw_evolve_cost", _bw_evolve_cost, "CUDA")


@torch.library.register_fake("qmps_b200::env_exact")
def _(A):
    N, D = A.shape[0], A.shape[-1]
    return (A.new_empty((N,)), A.new_empty((N, D, D)), A.new_empty((N, D, D)), A.new_empty((N,), dtype=torch.int32))


@torch.library.register_fake("qmps_b200::fixed_point_cost")
def _(A, B, outer):
    shape = (A.shape[0], B.shape[0]) if outer else (max(A.shape[0], B.shape[0]),)
    real = _RD[A.dtype]
    return (A.new_empty(shape), A.new_empty(shape, dtype=real), A.new_empty(shape, dtype=real), A.new_empty(shape, dtype=real))


@torch.library.register_fake("qmps_b200::tm_power")
def _(A, B, K):
    N, D = A.shape[0], A.shape[-1]
    return A.new_empty((N, D, D)), A.new_empty((N,))


@torch.library.register_fake("qmps_b200::bw_evolve_cost")
def _(U1, U2, V1, V2, W):
    return V1.new_empty((V1.reshape(-1, 4, 4).shape[0],), dtype=_RD[V1.dtype])
