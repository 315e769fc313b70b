"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares.
No compute call is made here."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "hybridq_b200" / "lib" / "libhybridq_b200.so"


def _declared():
    text = (ROOT / "include" / "hybridq_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([a-zA-Z_][a-zA-Z0-9_]*)\s*\(", text)
    skip = {"defined", "sizeof"}
    return sorted({n for n in names if n not in skip and (n.startswith("hq_") or n.startswith("swap_")
                   or n.startswith("apply_U") or n.startswith("to_complex") or n == "get_log2_pack_size")})


def _ensure_built():
    if not LIB.exists():
        import __graft_entry__ as g
        g.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    lib = ctypes.CDLL(str(LIB))
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hybridq_b200.h but not exported"


def test_reference_symbols_present_in_dropin_copies():
    _ensure_built()
    ref_syms = ["get_log2_pack_size", "apply_U_float32", "apply_U_float64", "to_complex64", "to_complex128"]
    swap_syms = [f"swap_{t}" for t in ("float32", "float64", "int32", "int64", "uint32", "uint64")]
    d = ROOT / "hybridq_b200" / "lib" / "dropin"
    u = ctypes.CDLL(str(d / "hybridq.so"))
    s = ctypes.CDLL(str(d / "hybridq_swap.so"))
    for n in ref_syms:
        assert hasattr(u, n)
    for n in swap_syms:
        assert hasattr(s, n)
    u.get_log2_pack_size.restype = ctypes.c_uint32
    assert u.get_log2_pack_size() >= 1        # 0 would read as "library missing" (simulation.py:393)


def test_python_binding_lists_every_symbol():
    _ensure_built()
    from hybridq_b200 import _lib
    assert sorted(_lib.EXPORTED) == _declared()


def test_abi_rejects_bad_arguments_without_touching_the_gpu():
    """Argument validation happens before any CUDA call (U.h:34-36, :48-54 contract)."""
    _ensure_built()
    import numpy as np
    from hybridq_b200 import _lib
    lib = _lib.lib
    n = 6
    buf = np.zeros(2 * 2 ** n + 16, dtype=np.float32)
    off = (-buf.ctypes.data // 4) % 8
    re = buf[off:off + 2 ** n]
    im = buf[off + 2 ** n:off + 2 * 2 ** n]
    U = np.eye(2, dtype=np.complex64)
    f = ctypes.POINTER(ctypes.c_float)
    u32 = ctypes.POINTER(ctypes.c_uint32)
    pos0 = np.array([0], dtype=np.uint32)          # below get_log2_pack_size() -> rc 1
    assert lib.apply_U_float32(re.ctypes.data_as(f), im.ctypes.data_as(f), U.ctypes.data_as(f),
                               pos0.ctypes.data_as(u32), n, 1) == 1
    pos = np.array([3], dtype=np.uint32)
    mis = buf[off + 1:off + 1 + 2 ** n]             # 4-byte aligned only -> rc 1
    assert lib.apply_U_float32(mis.ctypes.data_as(f), im.ctypes.data_as(f), U.ctypes.data_as(f),
                               pos.ctypes.data_as(u32), n, 1) == 1
    assert lib.apply_U_float32(re.ctypes.data_as(f), im.ctypes.data_as(f), U.ctypes.data_as(f),
                               pos.ctypes.data_as(u32), n, 0) == 0      # n_pos = 0 is a no-op
    dup = np.array([3, 3], dtype=np.uint32)
    assert lib.apply_U_float32(re.ctypes.data_as(f), im.ctypes.data_as(f), np.eye(4, dtype=np.complex64).ctypes.data_as(f),
                               dup.ctypes.data_as(u32), n, 2) == 1
    assert lib.swap_float32(re.ctypes.data_as(f), pos.ctypes.data_as(u32), n, 0) == 0
    bad = np.array([0, 0, 1], dtype=np.uint32)
    assert lib.swap_float32(re.ctypes.data_as(f), bad.ctypes.data_as(u32), n, 3) == 1
