"""oracle.py -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Python face of the CPU oracle for the state-vector evolution hot path:

* ``COracle``     ctypes wrapper over ``oracle/_build/libhq_oracle.so`` (the plain-C
                  restatement in ``hq_oracle.c`` of /root/reference/include/U.h:28-202,
                  swap.h:47-95, python_U.cpp:114-123).
* ``numpy_*``     an independent numpy restatement (tensordot / transpose), the same
                  idea as the reference's own ``dot(force_numpy=True)`` fallback
                  (/root/reference/hybridq/utils/dot.py:331-356).
* ``RefCore``     ctypes wrapper over the REFERENCE's own compiled core in
                  ``oracle/_ref/{wheel,avx2,native}`` with the exact prototypes the
                  reference binds (/root/reference/hybridq/utils/dot.py:49-71,
                  transpose.py:42-58).  Used to pin the oracle and as the CPU baseline.
* ``evolve_*``    gate-loop drivers over a list of ``(U, pos)`` gate-applies.
* ``numpy_project`` / ``numpy_measure``  numpy restatements of the reference's Projection and
                  Measure FunctionalGates (/root/reference/hybridq/gate/projection.py:25-116,
                  gate/measure.py:25-122) on a flat complex state.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  Parity status: PINNED (see
tests/test_oracle.py and the header of hq_oracle.c).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"
REFDIR = HERE / "_ref"

_c_float = {np.dtype("float32"): ctypes.c_float, np.dtype("float64"): ctypes.c_double}


def build(ref: bool = True) -> None:
    """Compile the C oracle (and oracle/_ref when /root/reference is present)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", str(HERE)] + targets, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def aligned_empty(shape, dtype, alignment: int = 128) -> np.ndarray:
    """Uninitialised C-contiguous array whose data pointer is `alignment`-aligned."""
    dtype = np.dtype(dtype)
    size = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(size + alignment, dtype=np.uint8)
    shift = (-raw.ctypes.data) % alignment
    return raw[shift:shift + size].view(dtype).reshape(shape)


def split_state(psi: np.ndarray, alignment: int = 128) -> np.ndarray:
    """complex (..) -> aligned real array (2, ..) = [re, im] planes."""
    psi = np.asarray(psi)
    ft = np.real(np.zeros(1, psi.dtype)).dtype
    out = aligned_empty((2,) + psi.shape, ft, alignment)
    np.copyto(out[0], psi.real)
    np.copyto(out[1], psi.imag)
    return out


def join_state(planes: np.ndarray) -> np.ndarray:
    return planes[0] + 1j * planes[1]


# ----------------------------------------------------------------------------------
# C oracle
# ----------------------------------------------------------------------------------
class COracle:
    def __init__(self, path: os.PathLike | None = None):
        path = Path(path) if path else BUILD / "libhq_oracle.so"
        if not path.exists():
            build(ref=False)
        self.lib = ctypes.CDLL(str(path))
        u32p = ctypes.POINTER(ctypes.c_uint32)
        for name, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p = ctypes.POINTER(ct)
            f = getattr(self.lib, f"oracle_apply_U_{name}")
            f.argtypes = [p, p, p, u32p, ctypes.c_uint, ctypes.c_uint]
            f.restype = ctypes.c_int
            g = getattr(self.lib, f"oracle_to_complex_{name}")
            g.argtypes = [p, p, p, ctypes.c_uint64]
            g.restype = ctypes.c_int
        for name in ("b32", "b64"):
            f = getattr(self.lib, f"oracle_swap_{name}")
            f.argtypes = [ctypes.c_void_p, u32p, ctypes.c_uint, ctypes.c_uint]
            f.restype = ctypes.c_int

    def apply_U(self, re: np.ndarray, im: np.ndarray, U: np.ndarray, pos) -> int:
        ft = re.dtype
        ct = _c_float[ft]
        ctype_c = np.dtype("complex64") if ft == np.float32 else np.dtype("complex128")
        U = np.ascontiguousarray(U, dtype=ctype_c)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        n = int(round(np.log2(re.size)))
        fn = self.lib.oracle_apply_U_f32 if ft == np.float32 else self.lib.oracle_apply_U_f64
        p = ctypes.POINTER(ct)
        return fn(re.ctypes.data_as(p), im.ctypes.data_as(p), U.ctypes.data_as(p),
                  pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n, len(pos))

    def swap(self, a: np.ndarray, pos) -> int:
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        n = int(round(np.log2(a.size)))
        fn = {4: self.lib.oracle_swap_b32, 8: self.lib.oracle_swap_b64}[a.dtype.itemsize]
        return fn(a.ctypes.data, pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n, len(pos))

    def to_complex(self, re: np.ndarray, im: np.ndarray) -> np.ndarray:
        ft = re.dtype
        out = np.empty(re.shape, dtype=np.complex64 if ft == np.float32 else np.complex128)
        p = ctypes.POINTER(_c_float[ft])
        fn = self.lib.oracle_to_complex_f32 if ft == np.float32 else self.lib.oracle_to_complex_f64
        fn(re.ctypes.data_as(p), im.ctypes.data_as(p), out.ctypes.data_as(p), re.size)
        return out


# ----------------------------------------------------------------------------------
# numpy restatement (independent second opinion; small n only)
# ----------------------------------------------------------------------------------
def numpy_apply_U(psi: np.ndarray, U: np.ndarray, pos) -> np.ndarray:
    """psi: complex flat (2^n).  Returns a new array.  pos[i] = index bit of matrix bit i."""
    n = int(round(np.log2(psi.size)))
    k = len(pos)
    t = psi.reshape((2,) * n)
    # matrix index bit k-1 is the most significant -> first tensor axis of U
    axes = [n - 1 - int(pos[i]) for i in reversed(range(k))]
    Ut = np.asarray(U).reshape((2,) * (2 * k))
    out = np.tensordot(Ut, t, axes=(list(range(k, 2 * k)), axes))
    out = np.moveaxis(out, list(range(k)), axes)
    return np.ascontiguousarray(out).reshape(-1)


def numpy_swap(a: np.ndarray, pos) -> np.ndarray:
    """new[j] = old[sigma(j)] on the low m bits; new bit i <- old bit pos[i]."""
    n = int(round(np.log2(a.size)))
    m = len(pos)
    t = a.reshape((2,) * n)
    # axis of bit b is n-1-b.  New axis for bit i takes old axis of bit pos[i].
    perm = list(range(n))
    for i in range(m):
        perm[n - 1 - i] = n - 1 - int(pos[i])
    return np.ascontiguousarray(np.transpose(t, perm)).reshape(-1)


# ----------------------------------------------------------------------------------
# the reference's own compiled core (oracle/_ref)
# ----------------------------------------------------------------------------------
class RefCore:
    """ctypes binding of the reference core with the reference's own prototypes."""

    def __init__(self, variant: str = "avx2", path: os.PathLike | None = None):
        """`path`: directory holding hybridq.so / hybridq_swap.so; defaults to oracle/_ref/<variant>.
        Any library exporting the reference's eleven symbols can be bound this way -- the GPU
        tests bind this repo's drop-in copies to drive them exactly like the reference does."""
        d = Path(path) if path is not None else REFDIR / variant
        self.variant = variant if path is None else str(path)
        self.lib_u = ctypes.CDLL(str(d / "hybridq.so"))
        self.lib_s = ctypes.CDLL(str(d / "hybridq_swap.so"))
        self.lib_u.get_log2_pack_size.restype = ctypes.c_uint32
        self.log2_pack_size = int(self.lib_u.get_log2_pack_size())
        u32p = ctypes.POINTER(ctypes.c_uint32)
        self._apply = {}
        self._swap = {}
        self._to_complex = {}
        for bits, ct in ((32, ctypes.c_float), (64, ctypes.c_double)):
            p = ctypes.POINTER(ct)
            f = getattr(self.lib_u, f"apply_U_float{bits}")
            f.argtypes = [p, p, p, u32p, ctypes.c_uint, ctypes.c_uint]
            f.restype = ctypes.c_int
            self._apply[np.dtype(f"float{bits}")] = f
            g = getattr(self.lib_u, f"to_complex{2 * bits}")
            g.argtypes = [p, p, p, ctypes.c_uint]
            g.restype = ctypes.c_int
            self._to_complex[np.dtype(f"float{bits}")] = g
        for t, ct in (("float32", ctypes.c_float), ("float64", ctypes.c_double),
                      ("int32", ctypes.c_int32), ("int64", ctypes.c_int64),
                      ("uint32", ctypes.c_uint32), ("uint64", ctypes.c_uint64)):
            f = getattr(self.lib_s, f"swap_{t}")
            f.argtypes = [ctypes.POINTER(ct), u32p, ctypes.c_uint, ctypes.c_uint]
            f.restype = ctypes.c_int
            self._swap[np.dtype(t)] = (f, ct)

    @staticmethod
    def available(variant: str) -> bool:
        d = REFDIR / variant
        return (d / "hybridq.so").exists() and (d / "hybridq_swap.so").exists()

    def apply_U(self, re, im, U, pos) -> int:
        ft = re.dtype
        p = ctypes.POINTER(_c_float[ft])
        U = np.ascontiguousarray(U, dtype=np.complex64 if ft == np.float32 else np.complex128)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        n = int(round(np.log2(re.size)))
        return self._apply[ft](re.ctypes.data_as(p), im.ctypes.data_as(p), U.ctypes.data_as(p),
                               pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n, len(pos))

    def swap(self, a, pos) -> int:
        f, ct = self._swap[a.dtype]
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        n = int(round(np.log2(a.size)))
        return f(a.ctypes.data_as(ctypes.POINTER(ct)),
                 pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n, len(pos))

    def to_complex(self, re, im) -> np.ndarray:
        ft = re.dtype
        out = np.empty(re.shape, dtype=np.complex64 if ft == np.float32 else np.complex128)
        p = ctypes.POINTER(_c_float[ft])
        self._to_complex[ft](re.ctypes.data_as(p), im.ctypes.data_as(p), out.ctypes.data_as(p),
                             re.size)
        return out


# ----------------------------------------------------------------------------------
# gate-loop drivers
# ----------------------------------------------------------------------------------
def numpy_project(psi: np.ndarray, pos, bits, renormalize: bool = True, atol: float = 1e-6) -> np.ndarray:
    """Projection of index bits `pos` onto the values `bits` (projection.py:70-116: the reference works on
    the re and im planes separately -- a plane whose projected norm is <= atol is dropped -- and then
    renormalises what is left)."""
    idx = np.arange(psi.size)
    keep = np.ones(psi.size, dtype=bool)
    for p, b in zip(pos, bits):
        keep &= ((idx >> int(p)) & 1) == int(b)
    re = np.where(keep, psi.real, 0)
    im = np.where(keep, psi.imag, 0)
    if not np.linalg.norm(re) > atol:
        re = np.zeros_like(re)
    if not np.linalg.norm(im) > atol:
        im = np.zeros_like(im)
    out = (re + 1j * im).astype(psi.dtype)
    if renormalize:
        norm = np.linalg.norm(out)
        if norm != 0:
            out = (out / norm).astype(psi.dtype)
    return out


def numpy_measure(psi: np.ndarray, pos_msb_first, renormalize: bool = True):
    """Measure (measure.py:25-75): probabilities of the outcomes of index bits `pos_msb_first` (first = most
    significant digit of the outcome), one draw from numpy's global generator, projection, renormalisation.
    Returns (new state, outcome)."""
    idx = np.arange(psi.size)
    k = len(pos_msb_first)
    outcome_of = np.zeros(psi.size, dtype=np.int64)
    for j, p in enumerate(pos_msb_first):
        outcome_of |= ((idx >> int(p)) & 1) << (k - 1 - j)
    probs = np.bincount(outcome_of, weights=(psi.real.astype(np.float64) ** 2 + psi.imag.astype(np.float64) ** 2),
                        minlength=2 ** k)
    probs = probs.astype(np.float32 if psi.dtype == np.complex64 else np.float64)
    s = int(np.random.choice(2 ** k, p=probs))
    out = np.where(outcome_of == s, psi, 0).astype(psi.dtype)
    if renormalize:
        out = (out / np.linalg.norm(out)).astype(psi.dtype)
    return out, s


def evolve_oracle(psi0: np.ndarray, gates, oracle: COracle | None = None) -> np.ndarray:
    """Apply `gates` = [(U, pos), ...] to complex psi0 (flat) with the C oracle."""
    oracle = oracle or COracle()
    planes = split_state(np.asarray(psi0).reshape(-1))
    for U, pos in gates:
        rc = oracle.apply_U(planes[0], planes[1], U, pos)
        if rc:
            raise RuntimeError(f"oracle_apply_U returned {rc}")
    return oracle.to_complex(planes[0], planes[1])


def evolve_numpy(psi0: np.ndarray, gates) -> np.ndarray:
    psi = np.array(psi0).reshape(-1)
    for U, pos in gates:
        psi = numpy_apply_U(psi, np.asarray(U, dtype=psi.dtype), pos)
    return psi


def evolve_ref(psi0: np.ndarray, gates, core: RefCore, timing: dict | None = None) -> np.ndarray:
    """Drive the reference core the way its own host loop does
    (/root/reference/hybridq/circuit/simulation/simulation.py:522-663): before a gate
    that touches an index bit below the core's pack width, permute the low bits so
    the targets sit above it, keep a bit map, and undo the permutation at the end.
    This is our own re-derivation of that bookkeeping (a running permutation
    `where[b]` = current physical bit of logical bit b), not a copy of the code."""
    import time
    psi0 = np.asarray(psi0).reshape(-1)
    n = int(round(np.log2(psi0.size)))
    L = core.log2_pack_size
    planes = split_state(psi0, alignment=128)
    re, im = planes[0], planes[1]
    where = list(range(n))          # logical bit -> physical bit
    t0 = time.perf_counter()
    for U, pos in gates:
        pos = [int(p) for p in pos]
        phys = [where[p] for p in pos]
        if any(p < L for p in phys):
            # window of low physical bits that is wide enough to park the targets above L
            k_low = None
            for w in range(L, n + 1):
                if sum(1 for p in phys if p < w) <= w - L:
                    k_low = w
                    break
            if k_low is None:
                raise RuntimeError("state too small for this gate with the reference core")
            inwin = [p for p in phys if p < k_low]
            others = [b for b in range(k_low) if b not in inwin]
            order = others + sorted(inwin)      # new physical bit i <- old physical bit order[i]
            if core.swap(re, order) or core.swap(im, order):
                raise RuntimeError("reference swap failed")
            newpos = {old: new for new, old in enumerate(order)}
            where = [newpos.get(w_, w_) for w_ in where]
            phys = [where[p] for p in pos]
        if core.apply_U(re, im, U, phys):
            raise RuntimeError("reference apply_U failed")
    # undo: want new physical bit b to hold logical bit b -> new bit b <- old bit where[b]
    m = max([b + 1 for b in range(n) if where[b] != b], default=0)
    if m:
        order = where[:m]
        if core.swap(re, order) or core.swap(im, order):
            raise RuntimeError("reference swap failed")
    if timing is not None:
        timing["gate_loop_s"] = time.perf_counter() - t0
    return core.to_complex(re, im)
