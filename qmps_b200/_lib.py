"""ctypes binding of the C ABI in ``include/qmps_b200.h``.

The shared library ``libqmps_b200.so`` is built in-tree by
``__graft_entry__.build()`` (nvcc, sm_100a).  There is NO fallback: if the library
is missing or no CUDA device is visible, every compute entry point raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqmps_b200.so")

C128, C64 = 0, 1
ST_OK, ST_NOT_PD, ST_NO_CONVERGE, ST_SINGULAR = 0, 1, 2, 3
GAUGE_TRACE, GAUGE_ZGEEV = 0, 1

# gate codes (include/qmps_b200.h)
G_RZ, G_RX, G_RY, G_H, G_CNOT, G_SWAP, G_CZ, G_XPOW, G_ZZPOW, G_XXPOW, G_YYPOW, G_X, G_Z = range(13)


class GateOp(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int32), ("q0", ctypes.c_int32), ("q1", ctypes.c_int32),
                ("param", ctypes.c_int32), ("scale", ctypes.c_double), ("offset", ctypes.c_double)]


class QmpsError(RuntimeError):
    pass


_vp, _i, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_gp = ctypes.POINTER(GateOp)

# name -> argtypes; every symbol include/qmps_b200.h declares
SIGNATURES = {
    "qmps_version": ([], ctypes.c_char_p),
    "qmps_last_error": ([], ctypes.c_char_p),
    "qmps_device_count": ([], _i),
    "qmps_set_option": ([ctypes.c_char_p, _i], _i),
    "qmps_debug_counters": ([_vp, _i], _i),
    "qmps_unitary_to_tensor": ([_i, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_tensor_to_unitary": ([_i, _i, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_environment_to_unitary": ([_i, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_env_exact": ([_i, _i, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_env_exact_host": ([_i, _i, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i], _i),
    "qmps_get_env_exact_host": ([_i, _i64, _vp, _vp, _vp, _i, _i], _i),
    "qmps_tm_apply": ([_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_scars_cost": ([_i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_scars_trajectory": ([_vp, _vp, _i, _i, _i, ctypes.c_double, ctypes.c_uint64, _i, _vp, _vp, _i, _vp], _i),
    "qmps_env_exact_packed": ([_i64, _vp, _i, _vp, _vp], _i),
    "qmps_env_exact_packed_host": ([_i64, _vp, _i, _vp, _i], _i),
    "qmps_fixed_point": ([_i, _i, _i64, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_fixed_point_ex": ([_i, _i, _i64, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_loschmidt_batched": ([_gp, _i, _i, _i64, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_loschmidt_batched_host": ([_gp, _i, _i, _i64, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i], _i),
    "qmps_energy_theta_host": ([_gp, _i, _i, _i64, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _i], _i),
    "qmps_rotosolve_sweep": ([_gp, _i, _i, _i64, _i, _vp, _vp, _i, _i, _vp, _i, _vp], _i),
    "qmps_argmin_allreduce": ([_vp, _i64, _vp, _i64, _vp, _vp, _vp], _i),
    "qmps_nccl_unique_id": ([_vp], _i),
    "qmps_nccl_comm_create": ([_vp, _i, _i, _vp], _i),
    "qmps_nccl_comm_destroy": ([_vp], _i),
    "qmps_tdvp_tangent": ([_i, _i, _i64, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_tdvp_dadt": ([_i, _i, _i64, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_tdvp_evolve": ([_i, _i, _i64, _vp, _vp, ctypes.c_double, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_loschmidt_trajectory": ([_gp, _i, _i, _i, _vp, _vp, _i, _i, _i, ctypes.c_double, ctypes.c_uint64, _i, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_merge": ([_i, _i, _i, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_ansatz": ([_gp, _i, _i, _i64, _i, _vp, _i, _vp, _i, _vp], _i),
    "qmps_energy_theta": ([_gp, _i, _i, _i64, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i, _vp], _i),
    "qmps_energy_tensor": ([_i, _i64, _vp, _i, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_rotosolve_fit": ([_i64, _i, _vp, _vp, _vp, _vp, _i, _i, _vp], _i),
    "qmps_tm_power": ([_i, _i, _i64, _vp, _vp, _vp, _i, _vp, _i, _vp], _i),
    "qmps_zgemm_c128_i8": ([_i64, _i, _i, _i, _vp, _vp, _i, _vp, _vp], _i),
    "qmps_cgemm_c64_tc": ([_i64, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp], _i),
    "qmps_left_canonicalise": ([_i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_mixed_canonical": ([_i, _i, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_gauge_transform": ([_i, _i, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_expectation": ([_i, _i, _i64, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp], _i),
    "qmps_bw_environment": ([_i, _i64, _i64, _vp, _vp, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_bw_env_apply": ([_i64, _i64, _vp, _vp, _i64, _vp, _vp, _i, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_bw_expectation": ([_i64, _i64, _vp, _vp, _i, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_bw_overlap": ([_i64, _i64, _vp, _vp, _i64, _vp, _vp, _i, _i64, _vp, _vp, _i64, _vp, _vp, _i, _vp], _i),
    "qmps_bw_evolve_cost": ([_i64, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp], _i),
    "qmps_argmin": ([_i64, _vp, _i64, _vp, _vp, _vp], _i),
    "qmps_loschmidt_rate": ([_i64, _vp, ctypes.c_double, ctypes.c_double, _vp, _vp], _i),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ``QmpsError`` if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QmpsError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  qmps_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def require_device():
    lib = load()
    if lib.qmps_device_count() < 1:
        raise QmpsError("qmps_b200 needs a CUDA device (sm_100a); none is visible and there is no CPU fallback")
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().qmps_last_error().decode("utf-8", "replace")
        raise QmpsError(f"{what} failed (code {rc}): {msg}")


def make_ops(op_list):
    """[(code, q0, q1, param, scale, offset), ...] -> ctypes array of GateOp."""
    arr = (GateOp * max(len(op_list), 1))()
    for k, (code, q0, q1, param, scale, offset) in enumerate(op_list):
        arr[k] = GateOp(int(code), int(q0), int(q1), int(param), float(scale), float(offset))
    return arr
