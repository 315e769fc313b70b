"""ctypes binding of libhybridq_b200.so (declared in include/hybridq_b200.h).

The library is loaded from ``hybridq_b200/lib/`` (built in-tree by
``__graft_entry__.build()`` / ``make -C hybridq_b200/csrc``).  There is NO fallback:
if the CUDA library is missing, importing this module raises -- the product path never
computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

LIBDIR = Path(__file__).resolve().parent / "lib"
LIBPATH = LIBDIR / "libhybridq_b200.so"
DROPIN_DIR = LIBDIR / "dropin"

C64, C128 = 0, 1


class HybridQB200Error(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    path = Path(os.environ.get("HYBRIDQ_B200_LIB", LIBPATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: the CUDA extension is not built.  Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C hybridq_b200/csrc`). "
            "hybridq_b200 has no CPU fallback.")
    return ctypes.CDLL(str(path))


lib = _load()

_u32p = ctypes.POINTER(ctypes.c_uint32)
_vp = ctypes.c_void_p


class PlanOptions(ctypes.Structure):
    _fields_ = [("tile_bits", ctypes.c_int), ("min_run_bits", ctypes.c_int), ("fuse", ctypes.c_int),
                ("max_gates_per_pass", ctypes.c_int), ("lookahead", ctypes.c_int),
                ("merge_max_k", ctypes.c_int), ("merge_pass_cost", ctypes.c_int),
                ("fast_slots", ctypes.c_int), ("mma_min_k", ctypes.c_int)]

    def __init__(self, tile_bits=0, min_run_bits=-1, fuse=1, max_gates_per_pass=0, lookahead=0,
                 merge_max_k=-1, merge_pass_cost=-1, fast_slots=1, mma_min_k=-1):
        super().__init__(tile_bits, min_run_bits, fuse, max_gates_per_pass, lookahead, merge_max_k,
                         merge_pass_cost, fast_slots, mma_min_k)


def _proto(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


# Part 1 (reference prototypes: hybridq/utils/dot.py:49-71, transpose.py:42-58)
_proto("get_log2_pack_size", ctypes.c_uint32)
for _b, _ct in ((32, ctypes.c_float), (64, ctypes.c_double)):
    _p = ctypes.POINTER(_ct)
    _proto(f"apply_U_float{_b}", ctypes.c_int, _p, _p, _p, _u32p, ctypes.c_uint, ctypes.c_uint)
    _proto(f"to_complex{2 * _b}", ctypes.c_int, _p, _p, _p, ctypes.c_uint)
for _t, _ct in (("float32", ctypes.c_float), ("float64", ctypes.c_double), ("int32", ctypes.c_int32),
                ("int64", ctypes.c_int64), ("uint32", ctypes.c_uint32), ("uint64", ctypes.c_uint64)):
    _proto(f"swap_{_t}", ctypes.c_int, ctypes.POINTER(_ct), _u32p, ctypes.c_uint, ctypes.c_uint)

# Part 2
_proto("hq_version", ctypes.c_int)
_proto("hq_last_error", ctypes.c_char_p)
_proto("hq_device_count", ctypes.c_int, ctypes.POINTER(ctypes.c_int))
_proto("hq_set_device", ctypes.c_int, ctypes.c_int)
_proto("hq_device_props", ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_size_t),
       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int))
_proto("hq_malloc", ctypes.c_int, ctypes.POINTER(_vp), ctypes.c_size_t)
_proto("hq_free", ctypes.c_int, _vp)
_proto("hq_host_alloc", ctypes.c_int, ctypes.POINTER(_vp), ctypes.c_size_t)
_proto("hq_host_free", ctypes.c_int, _vp)
_proto("hq_memcpy_h2d", ctypes.c_int, _vp, _vp, ctypes.c_size_t, _vp)
_proto("hq_memcpy_d2h", ctypes.c_int, _vp, _vp, ctypes.c_size_t, _vp)
_proto("hq_memcpy_d2d", ctypes.c_int, _vp, _vp, ctypes.c_size_t, _vp)
_proto("hq_stream_sync", ctypes.c_int, _vp)
_proto("hq_apply_U_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, _vp, _u32p, ctypes.c_uint, _vp)
_proto("hq_apply_U_direct_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, _vp, _u32p, ctypes.c_uint, _vp)
_proto("hq_swap_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, _u32p, ctypes.c_uint, _vp)
_proto("hq_pack_dev", ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, ctypes.c_uint64, _vp)
_proto("hq_unpack_dev", ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, ctypes.c_uint64, _vp)
_proto("hq_init_product_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, ctypes.c_char_p, _vp)
_proto("hq_init_random_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, ctypes.c_uint64, ctypes.c_uint64,
       ctypes.c_double, _vp)
_proto("hq_norm2_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), _vp)
_proto("hq_vdot_dev", ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), _vp)
_proto("hq_scale_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_double, _vp)
_proto("hq_marginal_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, _u32p, ctypes.c_uint,
       ctypes.POINTER(ctypes.c_double), _vp)
_proto("hq_project_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, _u32p, ctypes.c_uint, ctypes.c_uint,
       ctypes.c_double, ctypes.c_double, _vp)
_proto("hq_marginal_cond_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, _u32p, ctypes.c_uint, ctypes.c_uint64,
       ctypes.c_uint64, ctypes.POINTER(ctypes.c_double), _vp)
_proto("hq_project_mask_dev", ctypes.c_int, _vp, ctypes.c_int, ctypes.c_uint, ctypes.c_uint64, ctypes.c_uint64,
       ctypes.c_double, ctypes.c_double, _vp)
_proto("hq_plan_create", _vp, ctypes.c_int, ctypes.c_uint, ctypes.c_uint, _u32p, _u32p,
       ctypes.POINTER(ctypes.c_double), ctypes.POINTER(PlanOptions))
_proto("hq_plan_create_bitperm", _vp, ctypes.c_int, ctypes.c_uint, _u32p, ctypes.POINTER(PlanOptions))
_proto("hq_plan_destroy", None, _vp)
_proto("hq_plan_num_passes", ctypes.c_int, _vp)
_proto("hq_plan_num_gates", ctypes.c_int, _vp)
_proto("hq_plan_num_kernel_gates", ctypes.c_int, _vp)
_proto("hq_plan_flops", ctypes.c_double, _vp)
_proto("hq_plan_arith_counts", ctypes.c_int, _vp, _u32p, ctypes.c_int)
_proto("hq_plan_pass_info", ctypes.c_int, _vp, ctypes.c_int, _u32p, ctypes.c_int)
_proto("hq_plan_pass_gates", ctypes.c_int, _vp, ctypes.c_int, _u32p, ctypes.c_int)
_proto("hq_plan_run", ctypes.c_int, _vp, _vp, _vp)
_proto("hq_plan_run_range", ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp)
_proto("hq_plan_run_range_xchg", ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_uint,
       _u32p, ctypes.POINTER(_vp), _vp)
_proto("hq_plan_run_io", ctypes.c_int, _vp, _vp, _vp, _vp, _vp)
_proto("hq_host_is_pinned", ctypes.c_int, _vp)
_proto("hq_ipc_get_handle", ctypes.c_int, _vp, _vp)
_proto("hq_ipc_open", ctypes.c_int, _vp, ctypes.POINTER(_vp))
_proto("hq_ipc_close", ctypes.c_int, _vp)
_proto("hq_set_ring", ctypes.c_int, ctypes.c_int)
_proto("hq_set_umma", ctypes.c_int, ctypes.c_int)
_proto("hq_umma_launch_count", ctypes.c_uint64)
_proto("hq_direct_launch_count", ctypes.c_uint64)
_proto("hq_plan_umma_passes", ctypes.c_int, _vp)
_proto("hq_plan_sparse_rank_one_gates", ctypes.c_int, _vp)
_proto("hq_set_tuning", ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int)
_proto("hq_launch_count", ctypes.c_uint64)
_proto("hq_launch_count_reset", None)

EXPORTED = [
    "get_log2_pack_size", "apply_U_float32", "apply_U_float64", "to_complex64", "to_complex128",
    "swap_float32", "swap_float64", "swap_int32", "swap_int64", "swap_uint32", "swap_uint64",
    "hq_version", "hq_last_error", "hq_device_count", "hq_set_device", "hq_device_props", "hq_malloc",
    "hq_free", "hq_host_alloc", "hq_host_free", "hq_memcpy_h2d", "hq_memcpy_d2h", "hq_memcpy_d2d",
    "hq_stream_sync", "hq_apply_U_dev", "hq_apply_U_direct_dev", "hq_swap_dev", "hq_pack_dev",
    "hq_unpack_dev", "hq_init_product_dev", "hq_init_random_dev", "hq_norm2_dev", "hq_vdot_dev",
    "hq_scale_dev", "hq_marginal_dev", "hq_project_dev", "hq_marginal_cond_dev", "hq_project_mask_dev", "hq_plan_create", "hq_plan_create_bitperm", "hq_plan_destroy", "hq_plan_num_passes",
    "hq_plan_num_gates", "hq_plan_num_kernel_gates", "hq_plan_flops", "hq_plan_arith_counts", "hq_plan_pass_info", "hq_plan_pass_gates", "hq_plan_run", "hq_plan_run_range",
    "hq_plan_run_range_xchg", "hq_plan_run_io", "hq_host_is_pinned", "hq_ipc_get_handle", "hq_ipc_open", "hq_ipc_close", "hq_set_ring",
    "hq_set_umma", "hq_umma_launch_count", "hq_direct_launch_count", "hq_plan_umma_passes", "hq_plan_sparse_rank_one_gates",
    "hq_set_tuning", "hq_launch_count", "hq_launch_count_reset",
]


def last_error() -> str:
    return (lib.hq_last_error() or b"").decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise HybridQB200Error(f"{what or 'libhybridq_b200'} failed (rc={rc}): {last_error()}")


def dtype_code(complex_type) -> int:
    import numpy as np
    ct = np.dtype(complex_type)
    if ct == np.complex64:
        return C64
    if ct == np.complex128:
        return C128
    raise ValueError(f"unsupported complex type {ct}")
