"""Device-resident state vector and circuit plans (thin Python over the C ABI).

The state is ONE interleaved-complex array of 2^n amplitudes in HBM (a torch tensor is
used purely as the allocation / stream / NCCL handle; every operation on it goes through
libhybridq_b200.so with raw pointers).  This replaces the host-side split-plane buffer the
reference allocates per simulation (/root/reference/hybridq/circuit/simulation/simulation.py:491-509)
and makes its final ``to_complex`` pass (:669-675) unnecessary.
"""
from __future__ import annotations

import ctypes
import functools
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import lib, check, PlanOptions


def _stream_handle(stream=None, device=None) -> ctypes.c_void_p:
    """Raw cudaStream_t of `stream`, or of torch's current stream ON `device` (not on whatever device happens to
    be current: the C ABI launches on the calling thread's current device, see `_on`)."""
    import torch
    s = torch.cuda.current_stream(device) if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def _on(device):
    """Context manager making `device` the calling thread's current CUDA device for the duration of a C-ABI
    call: every hq_* entry point launches on cudaGetDevice()'s device, so a state that lives on another GPU
    than the current one must switch first (kernels ran in the wrong context otherwise)."""
    import torch
    return torch.cuda.device(device)


def _torch_ctype(complex_type):
    import torch
    return torch.complex64 if np.dtype(complex_type) == np.complex64 else torch.complex128


def _dev(method):
    """Run a DeviceState method with the state's device current (see `_on`)."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with _on(self.device):
            return method(self, *args, **kwargs)
    return wrapper


MARGINAL_MAX_K = 24      # HQ_MARGINAL_MAX_K: outcome bits per hq_marginal_cond_dev call


class Plan:
    """A circuit compiled into tile passes (hq_plan_*).  `gates` = [(U, pos), ...] with
    LSB-first index-bit positions, U a 2^k x 2^k array (row-major)."""

    def __init__(self, gates: Sequence, n_qubits: int, complex_type="complex64",
                 options: PlanOptions | None = None):
        self.n_qubits = int(n_qubits)
        self.complex_type = np.dtype(complex_type)
        self.dtype = _lib.dtype_code(complex_type)
        ks = np.array([len(p) for _, p in gates], dtype=np.uint32)
        pos = (np.concatenate([np.asarray(p, dtype=np.uint32).reshape(-1) for _, p in gates])
               if len(gates) else np.zeros(0, np.uint32))
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        mats = []
        for (U, p) in gates:
            U = np.asarray(U, dtype=np.complex128)
            if U.shape != (2 ** len(p), 2 ** len(p)):
                raise ValueError(f"matrix of shape {U.shape} does not match {len(p)} positions")
            mats.append(np.ascontiguousarray(U).reshape(-1))
        flat = (np.concatenate(mats) if mats else np.zeros(0, np.complex128)).view(np.float64)
        flat = np.ascontiguousarray(flat)
        self.options = options or PlanOptions()
        self._h = lib.hq_plan_create(self.dtype, self.n_qubits, len(gates),
                                     ks.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                     pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                     flat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                     ctypes.byref(self.options))
        if not self._h:
            raise _lib.HybridQB200Error(f"hq_plan_create failed: {_lib.last_error()}")
        self.n_gates = lib.hq_plan_num_gates(self._h)
        self.n_kernel_gates = lib.hq_plan_num_kernel_gates(self._h)
        self.flops = lib.hq_plan_flops(self._h)
        self.n_passes = lib.hq_plan_num_passes(self._h)
        # passes (one dense complex64 k = 4 .. 6 matrix) that run on the tcgen05 / TMEM kernel (hq_umma.cuh)
        self.n_umma_passes = lib.hq_plan_umma_passes(self._h)
        # scalar + rank-one gates (depolarizing channels) that run in the sparse form (hq_plan.cpp, fold of the scalars)
        self.n_sparse_rank_one = lib.hq_plan_sparse_rank_one_gates(self._h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.hq_plan_destroy(h)
            self._h = None

    def arithmetic(self) -> dict:
        """{'k2': {'ffma2_slots': .., 'tensor_cores': .., 'fma_generic': .., 'two_phase': ..}, ...}: how many
        kernel matrices of each size run on which arithmetic (hq_plan_arith_counts)."""
        out = (ctypes.c_uint32 * 40)()
        check(lib.hq_plan_arith_counts(self._h, out, 40), "hq_plan_arith_counts")
        names = ("ffma2_slots", "tensor_cores", "fma_generic", "two_phase", "scalar_plus_rank_one")
        res = {}
        for k in range(1, 9):
            row = {names[a]: int(out[5 * (k - 1) + a]) for a in range(5) if out[5 * (k - 1) + a]}
            if row:
                res[f"k{k}"] = row
        return res

    def pass_info(self, p: int) -> dict:
        out = (ctypes.c_uint32 * 32)()
        check(lib.hq_plan_pass_info(self._h, p, out, 32), "hq_plan_pass_info")
        ng = out[4]
        ids = (ctypes.c_uint32 * max(1, ng))()
        check(lib.hq_plan_pass_gates(self._h, p, ids, max(1, ng)), "hq_plan_pass_gates")
        return {"tile_bits": out[0], "n_high": out[1], "n_gates": ng, "n_kernel_gates": out[2],
                "has_perm": out[3], "high_pos": [out[5 + i] for i in range(out[1])],
                "fast_mask": out[5 + out[1]], "chain_mask": out[6 + out[1]],
                "gate_ids": [ids[i] for i in range(ng)]}

    def run(self, state: "DeviceState", first: int | None = None, last: int | None = None, stream=None):
        if state.n_qubits != self.n_qubits or state.dtype != self.dtype:
            raise ValueError("plan and state disagree on size or precision")
        with _on(state.device):
            if first is None and last is None:
                check(lib.hq_plan_run(self._h, state.ptr, _stream_handle(stream, state.device)), "hq_plan_run")
            else:
                check(lib.hq_plan_run_range(self._h, state.ptr, first or 0,
                                            self.n_passes if last is None else last,
                                            _stream_handle(stream, state.device)), "hq_plan_run_range")


    def run_io(self, state: "DeviceState", host_src: np.ndarray | None, host_dst: np.ndarray | None, stream=None):
        """All passes with the upload and / or download folded in (hq_plan_run_io): the first pass reads its tiles
        from the PINNED host array `host_src`, the last pass writes the result to the PINNED host array `host_dst`.
        Asynchronous; call state.sync() before reading `host_dst`."""
        if state.n_qubits != self.n_qubits or state.dtype != self.dtype:
            raise ValueError("plan and state disagree on size or precision")
        for a in (host_src, host_dst):
            if a is not None and (a.size != state.n_amps or a.dtype != state.complex_type or not a.flags.c_contiguous):
                raise ValueError("host arrays must be contiguous, of the state's size and precision")
        with _on(state.device):
            check(lib.hq_plan_run_io(self._h, state.ptr,
                                     ctypes.c_void_p(host_src.ctypes.data) if host_src is not None else None,
                                     ctypes.c_void_p(host_dst.ctypes.data) if host_dst is not None else None,
                                     _stream_handle(stream, state.device)), "hq_plan_run_io")

    def run_xchg(self, state: "DeviceState", gbit_digit: int, lpos: Sequence[int], dst_ptrs: Sequence[int], stream=None):
        """All passes, the last one with its write-back redirected (hq_plan_run_range_xchg): the amplitudes whose
        local index bits `lpos` spell D go to the buffer `dst_ptrs[D]` at the same index with those bits replaced
        by `gbit_digit`.  A plan without passes becomes one gate-less redirect pass."""
        if state.n_qubits != self.n_qubits or state.dtype != self.dtype:
            raise ValueError("plan and state disagree on size or precision")
        s = len(lpos)
        pos = (ctypes.c_uint32 * 4)(*(list(lpos) + [0] * (4 - s)))
        dst = (ctypes.c_void_p * 8)(*([int(p) for p in dst_ptrs] + [None] * (8 - len(dst_ptrs))))
        with _on(state.device):
            check(lib.hq_plan_run_range_xchg(self._h, state.ptr, 0, self.n_passes, s, int(gbit_digit), pos, dst,
                                             _stream_handle(stream, state.device)), "hq_plan_run_range_xchg")


class BitPermPlan:
    """In-place permutation of index bits as tile passes (hq_plan_create_bitperm):
    new index bit i <- old index bit perm[i]."""

    def __init__(self, perm: Sequence[int], n_qubits: int, complex_type="complex64",
                 options: PlanOptions | None = None):
        self.n_qubits = int(n_qubits)
        self.dtype = _lib.dtype_code(complex_type)
        perm = np.ascontiguousarray(perm, dtype=np.uint32)
        if len(perm) != self.n_qubits:
            raise ValueError("perm must list every bit")
        self._h = lib.hq_plan_create_bitperm(self.dtype, self.n_qubits,
                                             perm.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
                                             ctypes.byref(options or PlanOptions()))
        if not self._h:
            raise _lib.HybridQB200Error(f"hq_plan_create_bitperm failed: {_lib.last_error()}")
        self.n_passes = lib.hq_plan_num_passes(self._h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.hq_plan_destroy(h)
            self._h = None

    def run(self, state: "DeviceState", stream=None):
        if state.n_qubits != self.n_qubits or state.dtype != self.dtype:
            raise ValueError("plan and state disagree on size or precision")
        with _on(state.device):
            check(lib.hq_plan_run(self._h, state.ptr, _stream_handle(stream, state.device)), "hq_plan_run")


class RawDeviceBuffer:
    """A cudaMalloc'ed block owned by this object (hq_malloc / hq_free), exposed to torch through
    ``__cuda_array_interface__``.  Shard buffers that peers map over NVLink are allocated this way: a
    cudaIpc handle needs the BASE pointer of an allocation, which a block carved out of torch's caching
    allocator is not."""

    def __init__(self, nbytes: int, device: int):
        self.device = int(device)
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        with _on(self.device):
            check(lib.hq_malloc(ctypes.byref(p), self.nbytes), "hq_malloc")
        self.ptr = p.value
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self, complex_type):
        import torch
        t = torch.as_tensor(self, device=f"cuda:{self.device}")
        return t.view(_torch_ctype(complex_type))

    def ipc_handle(self) -> bytes:
        h = ctypes.create_string_buffer(64)
        with _on(self.device):
            check(lib.hq_ipc_get_handle(ctypes.c_void_p(self.ptr), h), "hq_ipc_get_handle")
        return h.raw

    def __del__(self):
        p, self.ptr = getattr(self, "ptr", None), None
        if p:
            try:
                with _on(self.device):
                    lib.hq_free(ctypes.c_void_p(p))
            except Exception:
                pass


def open_ipc(handle: bytes, device: int) -> int:
    """Map a peer's RawDeviceBuffer into this process for kernels running on `device` (cudaIpcOpenMemHandle with
    lazy peer access); returns the device pointer."""
    p = ctypes.c_void_p()
    with _on(device):
        check(lib.hq_ipc_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(p)), "hq_ipc_open")
    return p.value


class DeviceState:
    def __init__(self, n_qubits: int, complex_type="complex64", device: int | None = None, tensor=None,
                 ipc: bool = False):
        import torch
        if not torch.cuda.is_available():
            raise _lib.HybridQB200Error("hybridq_b200 needs a CUDA device (no CPU fallback)")
        self.n_qubits = int(n_qubits)
        self.complex_type = np.dtype(complex_type)
        self.dtype = _lib.dtype_code(complex_type)
        if device is None:
            device = torch.cuda.current_device()
        self.device = int(device)
        self.raw = None
        if tensor is None and ipc:
            # own cudaMalloc block, so that peers can map it (see RawDeviceBuffer)
            self.raw = RawDeviceBuffer((2 ** self.n_qubits) * self.complex_type.itemsize, self.device)
            tensor = self.raw.tensor(self.complex_type)
        if tensor is None:
            with torch.cuda.device(self.device):
                tensor = torch.empty(2 ** self.n_qubits, dtype=_torch_ctype(complex_type),
                                     device=f"cuda:{self.device}")
        self.tensor = tensor
        self.n_amps = 2 ** self.n_qubits
        self.nbytes = self.n_amps * self.complex_type.itemsize

    @property
    def ptr(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(self.tensor.data_ptr())

    # -- transfers ------------------------------------------------------------------
    @_dev
    def upload(self, psi: np.ndarray, stream=None, sync: bool = True) -> "DeviceState":
        psi = np.asarray(psi)
        if psi.size != self.n_amps:
            raise ValueError("wrong number of amplitudes")
        if psi.dtype != self.complex_type or not psi.flags.c_contiguous:
            psi = np.ascontiguousarray(psi, dtype=self.complex_type)
        s = _stream_handle(stream, self.device)
        check(lib.hq_memcpy_h2d(self.ptr, ctypes.c_void_p(psi.ctypes.data), self.nbytes, s), "h2d")
        if sync:
            check(lib.hq_stream_sync(s), "sync")
        return self

    @_dev
    def download(self, out: np.ndarray | None = None, stream=None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n_amps, dtype=self.complex_type)
        if out.size != self.n_amps or out.dtype != self.complex_type or not out.flags.c_contiguous:
            raise ValueError("bad output array")
        s = _stream_handle(stream, self.device)
        check(lib.hq_memcpy_d2h(ctypes.c_void_p(out.ctypes.data), self.ptr, self.nbytes, s), "d2h")
        check(lib.hq_stream_sync(s), "sync")
        return out

    @_dev
    def sync(self, stream=None) -> None:
        check(lib.hq_stream_sync(_stream_handle(stream, self.device)), "sync")

    # -- preparation / reductions -----------------------------------------------------
    @_dev
    def init_product(self, spec: str, stream=None) -> "DeviceState":
        if len(spec) == 1:
            spec = spec * self.n_qubits
        check(lib.hq_init_product_dev(self.ptr, self.dtype, self.n_qubits, spec.encode(),
                                      _stream_handle(stream, self.device)), "hq_init_product_dev")
        return self

    @_dev
    def init_random(self, seed: int = 0, index_offset: int = 0, scale: float = 0.0, stream=None):
        check(lib.hq_init_random_dev(self.ptr, self.dtype, self.n_qubits, seed, index_offset, scale,
                                     _stream_handle(stream, self.device)), "hq_init_random_dev")
        return self

    @_dev
    def norm2(self, stream=None) -> float:
        r = ctypes.c_double()
        check(lib.hq_norm2_dev(self.ptr, self.dtype, self.n_amps, ctypes.byref(r), _stream_handle(stream, self.device)),
              "hq_norm2_dev")
        return r.value

    @_dev
    def vdot(self, other: "DeviceState", stream=None) -> complex:
        """<self|other> = sum conj(self) * other."""
        r = (ctypes.c_double * 2)()
        check(lib.hq_vdot_dev(self.ptr, other.ptr, self.dtype, self.n_amps, r, _stream_handle(stream, self.device)),
              "hq_vdot_dev")
        return complex(r[0], r[1])

    @_dev
    def marginal(self, pos: Sequence[int], stream=None, cond_mask: int = 0, cond_value: int = 0) -> np.ndarray:
        """(2^k, 2) array: sum of re^2 and of im^2 over the amplitudes whose index bits ``pos`` spell the
        outcome s (bit j of s = index bit pos[j]), restricted to the indices with
        ``(index & cond_mask) == cond_value``.  ``.sum(axis=1)`` are the measurement probabilities of
        the reference's ``_Measure(get_probs_only=True)`` (hybridq/gate/measure.py:25-50).  k <= 24 per call."""
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        if len(pos) > MARGINAL_MAX_K:
            raise ValueError(f"at most {MARGINAL_MAX_K} outcome bits per call (condition on earlier chunks)")
        out = np.zeros((2 ** len(pos), 2), dtype=np.float64)
        check(lib.hq_marginal_cond_dev(self.ptr, self.dtype, self.n_qubits,
                                       pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(pos),
                                       int(cond_mask), int(cond_value),
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                       _stream_handle(stream, self.device)), "hq_marginal_cond_dev")
        return out

    @_dev
    def project(self, pos: Sequence[int], outcome: int, scale_re: float = 1.0, scale_im: float = 1.0, stream=None):
        """Zero every amplitude whose index bits ``pos`` do not spell ``outcome``; scale the others
        plane-wise (hybridq/gate/projection.py:25-68)."""
        mask = value = 0
        for j, p in enumerate(pos):
            mask |= 1 << int(p)
            value |= ((int(outcome) >> j) & 1) << int(p)
        check(lib.hq_project_mask_dev(self.ptr, self.dtype, self.n_qubits, mask, value,
                                      float(scale_re), float(scale_im), _stream_handle(stream, self.device)),
              "hq_project_mask_dev")
        return self

    @_dev
    def scale(self, factor: float, stream=None):
        check(lib.hq_scale_dev(self.ptr, self.dtype, self.n_amps, float(factor), _stream_handle(stream, self.device)),
              "hq_scale_dev")
        return self

    # -- checkpoint / sampling ----------------------------------------------------------
    @_dev
    def dump(self, path, chunk_bytes: int = 1 << 28, meta: dict | None = None) -> None:
        """Write the state to `path` (raw little-endian interleaved complex, preceded by nothing) plus
        `path + '.json'` describing it; the copy goes through a pinned staging buffer chunk by chunk, so a state
        larger than host RAM can be checkpointed.  The reference has no checkpointing."""
        import json
        import torch
        from pathlib import Path
        path = Path(path)
        n_chunk = max(1, min(self.nbytes, int(chunk_bytes)) // self.complex_type.itemsize)
        stage = torch.empty(n_chunk, dtype=_torch_ctype(self.complex_type), pin_memory=True)
        host = stage.numpy()
        s = _stream_handle(None, self.device)
        with open(path, "wb") as f, _on(self.device):
            for a in range(0, self.n_amps, n_chunk):
                m = min(n_chunk, self.n_amps - a)
                check(lib.hq_memcpy_d2h(ctypes.c_void_p(host.ctypes.data),
                                        ctypes.c_void_p(self.tensor.data_ptr() + a * self.complex_type.itemsize),
                                        m * self.complex_type.itemsize, s), "d2h")
                check(lib.hq_stream_sync(s), "sync")
                f.write(memoryview(host[:m]))
        info = {"format": "hybridq_b200 state v1", "n_qubits": self.n_qubits, "complex_type": str(self.complex_type),
                "n_amps": self.n_amps, "layout": "interleaved complex, index bit 0 = LSB, first sorted qubit = MSB"}
        info.update(meta or {})
        Path(str(path) + ".json").write_text(json.dumps(info))

    @_dev
    def load(self, path, chunk_bytes: int = 1 << 28, expect: dict | None = None) -> "DeviceState":
        """Read a state written by :meth:`dump` (sizes, precision and any `expect`ed metadata must match)."""
        import json
        import torch
        from pathlib import Path
        path = Path(path)
        info = json.loads(Path(str(path) + ".json").read_text())
        want = {"n_qubits": self.n_qubits, "complex_type": str(self.complex_type)}
        want.update(expect or {})
        for k, v in want.items():
            if info.get(k) != v:
                raise ValueError(f"checkpoint {path}: {k} = {info.get(k)!r}, expected {v!r}")
        if path.stat().st_size != self.nbytes:
            raise ValueError(f"checkpoint {path}: {path.stat().st_size} bytes, expected {self.nbytes}")
        n_chunk = max(1, min(self.nbytes, int(chunk_bytes)) // self.complex_type.itemsize)
        stage = torch.empty(n_chunk, dtype=_torch_ctype(self.complex_type), pin_memory=True)
        host = stage.numpy()
        s = _stream_handle(None, self.device)
        with open(path, "rb") as f, _on(self.device):
            for a in range(0, self.n_amps, n_chunk):
                m = min(n_chunk, self.n_amps - a)
                f.readinto(memoryview(host[:m]).cast("B"))
                check(lib.hq_memcpy_h2d(ctypes.c_void_p(self.tensor.data_ptr() + a * self.complex_type.itemsize),
                                        ctypes.c_void_p(host.ctypes.data), m * self.complex_type.itemsize, s), "h2d")
                check(lib.hq_stream_sync(s), "sync")
        return self

    @_dev
    def sample(self, n_samples: int, seed=None, block_bits: int = 16) -> np.ndarray:
        """Draw `n_samples` basis-state indices from |psi|^2 WITHOUT collapsing the state (int64 array; bit b of
        an index = index bit b, i.e. `format(i, f'0{n}b')` reads first sorted qubit first).  Two levels: one
        device reduction gives the probability of every block of 2^block_bits consecutive amplitudes (the
        marginal of the high index bits), the block of each sample is drawn on the host, and only the drawn
        blocks are copied back to finish the draw inside them."""
        rng = np.random.default_rng(seed)
        n = self.n_qubits
        B = min(n, max(int(block_bits), n - MARGINAL_MAX_K))
        hi = list(range(B, n))
        p_blocks = self.marginal(hi).sum(axis=1) if hi else np.array([self.norm2()])
        p_blocks = p_blocks / p_blocks.sum()
        blocks = rng.choice(len(p_blocks), size=int(n_samples), p=p_blocks)
        out = np.empty(int(n_samples), dtype=np.int64)
        buf = np.empty(2 ** B, dtype=self.complex_type)
        s = _stream_handle(None, self.device)
        for blk in np.unique(blocks):
            check(lib.hq_memcpy_d2h(ctypes.c_void_p(buf.ctypes.data),
                                    ctypes.c_void_p(self.tensor.data_ptr() + int(blk) * (2 ** B) * self.complex_type.itemsize),
                                    buf.nbytes, s), "d2h")
            check(lib.hq_stream_sync(s), "sync")
            p = buf.real.astype(np.float64) ** 2 + buf.imag.astype(np.float64) ** 2
            sel = np.flatnonzero(blocks == blk)
            out[sel] = (int(blk) << B) | rng.choice(2 ** B, size=sel.size, p=p / p.sum())
        return out

    def copy(self) -> "DeviceState":
        return DeviceState(self.n_qubits, self.complex_type, self.device, tensor=self.tensor.clone())

    # -- gates -------------------------------------------------------------------------
    @_dev
    def apply(self, U: np.ndarray, pos: Sequence[int], stream=None, direct: bool = False):
        """One gate-apply (hq_apply_U_dev): pos[i] = index bit of matrix bit i, any bits."""
        U = np.ascontiguousarray(U, dtype=self.complex_type)
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        fn = lib.hq_apply_U_direct_dev if direct else lib.hq_apply_U_dev
        check(fn(self.ptr, self.dtype, self.n_qubits, ctypes.c_void_p(U.ctypes.data),
                 pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(pos), _stream_handle(stream, self.device)),
              "hq_apply_U_dev")
        return self

    @_dev
    def swap(self, pos: Sequence[int], stream=None):
        """In-place permutation of the low len(pos) index bits (hq_swap_dev)."""
        pos = np.ascontiguousarray(pos, dtype=np.uint32)
        check(lib.hq_swap_dev(self.ptr, self.dtype, self.n_qubits,
                              pos.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), len(pos),
                              _stream_handle(stream, self.device)), "hq_swap_dev")
        return self

    @_dev
    def permute_bits(self, perm: Sequence[int], options: PlanOptions | None = None, stream=None):
        """new index bit i <- old index bit perm[i] for every i < n (in place)."""
        BitPermPlan(perm, self.n_qubits, self.complex_type, options).run(self, stream)
        self.sync(stream)
        return self
