"""User-level ``dot`` / ``transpose`` / ``to_complex`` -- mirrors of the reference wrappers
over the native core (/root/reference/hybridq/utils/dot.py:139 ``dot``, :101 ``to_complex``,
/root/reference/hybridq/utils/transpose.py:61 ``transpose``), running on the GPU.

Same arguments and the same validation errors as the reference.  Two deliberate
differences, both consequences of the kernels taking any target bit:

* no low-bit swap window is ever needed (dot.py:217-221, :278-299), so ``swap_back=False``
  returns ``(result, None)``;
* shapes the native core cannot take (non-qubit dimensions, non-square matrix) raise
  ``NotImplementedError`` instead of silently switching to ``numpy.dot`` -- this package
  has no CPU compute path.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import lib, check
from .state import DeviceState

_SWAP = {np.dtype(t): (getattr(lib, f"swap_{t}"), ct) for t, ct in (
    ("float32", ctypes.c_float), ("float64", ctypes.c_double), ("int32", ctypes.c_int32),
    ("int64", ctypes.c_int64), ("uint32", ctypes.c_uint32), ("uint64", ctypes.c_uint64))}


def to_complex(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """re/im planes -> complex array through to_complex64/128 (dot.py:101-136)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        raise ValueError("'a' and 'b' must have the same shape.")
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        raise ValueError("Both 'a' and 'b' must be real valued.")
    if a.dtype != b.dtype or a.dtype not in (np.dtype("float32"), np.dtype("float64")):
        raise NotImplementedError("to_complex needs two float32 or two float64 arrays")
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    ct = ctypes.c_float if a.dtype == np.float32 else ctypes.c_double
    out = np.empty(a.shape, dtype=np.complex64 if a.dtype == np.float32 else np.complex128)
    fn = lib.to_complex64 if a.dtype == np.float32 else lib.to_complex128
    if a.size >= 2 ** 32:
        raise ValueError("to_complex64/128 take a 32-bit size (python_U.cpp:115-116)")
    p = ctypes.POINTER(ct)
    check(fn(a.ctypes.data_as(p), b.ctypes.data_as(p), out.ctypes.data_as(p), a.size), "to_complex")
    return out


def dot(a: np.ndarray, b: np.ndarray, axes_b=None, b_as_complex_array: bool = False,
        inplace: bool = False, backend="numpy", **kwargs):
    """Apply the square matrix `a` to the axes `axes_b` of the all-dimensions-2 array `b`."""
    if backend != "numpy":
        raise ValueError(f"Backend {backend} is not supported.")
    kwargs.setdefault("swap_back", True)
    kwargs.setdefault("force_numpy", False)
    kwargs.setdefault("raise_if_hcore_fails", False)
    if kwargs["force_numpy"]:
        raise NotImplementedError("hybridq_b200 has no numpy compute path")
    if axes_b is None:
        raise NotImplementedError("plain matrix products are not part of the evolution core")

    a = np.asarray(a, order="C")
    b_in = b
    b = np.asarray(b, order="C")
    b_ndim = b.ndim - (1 if b_as_complex_array else 0)
    b_shape = np.asarray(b.shape[1:] if b_as_complex_array else b.shape)
    axes_b = np.asarray(axes_b)
    if b_as_complex_array:
        if b.shape[0] != 2:
            raise ValueError("'b' is in the wrong format.")
        if np.iscomplexobj(b):
            raise ValueError("'b' is expected to be real.")
    real_type = b.dtype if b_as_complex_array else np.real(np.array([1], dtype=b.dtype)).dtype
    if real_type not in (np.dtype("float32"), np.dtype("float64")):
        raise NotImplementedError(f"unsupported type {real_type}")
    complex_type = np.dtype("complex64") if real_type == np.float32 else np.dtype("complex128")
    if any(axes_b >= b_ndim):
        raise IndexError("Index not in 'b'")
    if a.shape[-1] != np.prod(b_shape[axes_b]):
        raise ValueError("'a' and 'b' are incompatible.")
    if not (a.ndim == 2 and a.shape[0] == a.shape[1] and all(x == 2 for x in b_shape)
            and len(axes_b) <= 10):
        raise NotImplementedError("the native core needs a square matrix and a (2,)*n array")

    pos = (b_ndim - axes_b[::-1] - 1).astype("uint32")          # dot.py:215
    if b_as_complex_array:
        psi = np.empty(2 ** b_ndim, dtype=complex_type)
        psi.real = b[0].reshape(-1)
        psi.imag = b[1].reshape(-1)
    else:
        psi = np.ascontiguousarray(b.reshape(-1), dtype=complex_type)
    state = DeviceState(b_ndim, complex_type).upload(psi)
    state.apply(np.asarray(a, dtype=complex_type), pos)
    out = state.download()

    if b_as_complex_array:
        res = b_in if (inplace and isinstance(b_in, np.ndarray) and b_in.flags.c_contiguous
                       and b_in.flags.writeable) else np.empty_like(b)
        res[0] = out.real.reshape(b.shape[1:])
        res[1] = out.imag.reshape(b.shape[1:])
    else:
        res = out.reshape(b.shape)
        if inplace and isinstance(b_in, np.ndarray) and b_in.dtype == complex_type and b_in.flags.writeable:
            b_in[...] = res
            res = b_in
    return res if kwargs["swap_back"] is True else (res, None)


def transpose(a: np.ndarray, axes=None, inplace: bool = False, backend="numpy", **kwargs):
    """Transpose an all-dimensions-2 array through the swap_* core (transpose.py:61-160)."""
    if backend != "numpy":
        raise ValueError(f"Backend {backend} is not supported.")
    kwargs.setdefault("force_numpy", False)
    if kwargs["force_numpy"]:
        raise NotImplementedError("hybridq_b200 has no numpy compute path")
    a_in = a
    a = np.asarray(a, order="C")
    if axes is None:
        axes = np.arange(a.ndim).astype(np.uint32)[::-1]
    else:
        axes = np.asarray(axes, dtype=np.uint32)
    if len(axes) != a.ndim or any(x >= a.ndim for x in axes):
        raise IndexError("'axes' out of range.")
    if a.dtype not in _SWAP or a.shape != (2,) * a.ndim:
        raise NotImplementedError("the native core needs a (2,)*n array of a 4- or 8-byte real type")
    n_ord = next((i for i, x in enumerate(axes) if i != x), len(axes))
    if n_ord == len(axes):
        return a
    if not (inplace and a is a_in and a.flags.writeable):
        a = np.array(a)
    tail = axes[n_ord:]
    pos = np.ascontiguousarray(a.ndim - tail[::-1] - 1, dtype=np.uint32)     # transpose.py:139
    fn, ct = _SWAP[a.dtype]
    check(fn(a.ctypes.data_as(ctypes.POINTER(ct)), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
             a.ndim, len(pos)), "swap")
    return a
