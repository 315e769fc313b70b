"""Shared helpers for the tests (test infrastructure; may use the oracle)."""
import ctypes
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
TOL = {"complex64": 1e-6, "complex128": 1e-12}     # max-abs amplitude error (BASELINE.json north_star)


def golden_gates(z, prefix):
    """[(U, pos)] of a stored circuit; stored qubit indices follow the sorted-qubit order, so
    index i maps to bit n-1-i and gate.qubits is MSB-first (simulation.py:512-513, :633)."""
    n = int(z[f"{prefix}_nqubits"]) if f"{prefix}_nqubits" in z else None
    ng = int(z[f"{prefix}_ngates"])
    gates = []
    for j in range(ng):
        U = z[f"{prefix}_g{j}_U"]
        q = z[f"{prefix}_g{j}_q"]
        gates.append((U, q))
    return gates, n


def lower(gates_q, n):
    return [(U, [n - 1 - int(x) for x in reversed(q)]) for U, q in gates_q]


def functional_items(z, i):
    """Items of a stored circuit with FunctionalGates: ('U', matrix, qubits) | ('P', '01..', qubits) | ('M', None, qubits)."""
    items = []
    for j in range(int(z[f"f{i}_nitems"])):
        kind = str(z[f"f{i}_i{j}_kind"])
        q = [int(x) for x in z[f"f{i}_i{j}_q"]]
        payload = z[f"f{i}_i{j}_U"] if kind == "U" else (str(z[f"f{i}_i{j}_state"]) if kind == "P" else None)
        items.append((kind, payload, q))
    return items


def initial_from(z, key, n, ctype):
    init = z[key]
    if init.dtype.kind in "US":
        return str(init)
    return np.asarray(init, dtype=ctype).reshape(-1)


def product_state(spec, ctype):
    """numpy restatement of prepare_state for strings of 0,1,+,- (simulation/utils.py:40-156)."""
    v = {"0": [1, 0], "1": [0, 1], "+": [2 ** -0.5, 2 ** -0.5], "-": [2 ** -0.5, -(2 ** -0.5)]}
    psi = np.array([1.0])
    for c in spec:
        psi = np.kron(psi, np.array(v[c]))
    return psi.astype(ctype)


_EMU_DEFAULTS = (0, -1, 1, 0, 0, -1, -1, 1, -1)   # PlanOptions field order, mma_min_k last


def _emu_opts(opts):
    if not opts:
        return None
    o = tuple(opts)
    return (ctypes.c_int * 9)(*(o + _EMU_DEFAULTS[len(o):]))


class Emu:
    """ctypes face of libhq_emu.so (CPU model of the kernel phases + the planner)."""

    def __init__(self):
        path = ROOT / "hybridq_b200" / "lib" / "libhq_emu.so"
        if not path.exists():
            import subprocess
            subprocess.run(["make", "-s", "-C", str(ROOT / "hybridq_b200" / "csrc"), "emu"], check=True)
        self.lib = ctypes.CDLL(str(path))

    @staticmethod
    def _pack(gates):
        ks = np.array([len(p) for _, p in gates], dtype=np.uint32)
        pos = (np.concatenate([np.asarray(p, dtype=np.uint32) for _, p in gates])
               if gates else np.zeros(0, np.uint32)).astype(np.uint32)
        U = (np.concatenate([np.asarray(u, dtype=np.complex128).reshape(-1) for u, _ in gates])
             if gates else np.zeros(0, np.complex128)).view(np.float64)
        return ks, np.ascontiguousarray(pos), np.ascontiguousarray(U)

    def run(self, psi, gates, opts=None):
        dt = 0 if psi.dtype == np.complex64 else 1
        n = int(round(np.log2(psi.size)))
        ks, pos, U = self._pack(gates)
        st = np.ascontiguousarray(psi).copy()
        o = _emu_opts(opts)
        info = (ctypes.c_int * 4)()
        err = ctypes.create_string_buffer(256)
        rc = self.lib.hq_emu_run_circuit(dt, n, len(gates), ks.ctypes.data_as(ctypes.c_void_p),
                                         pos.ctypes.data_as(ctypes.c_void_p), U.ctypes.data_as(ctypes.c_void_p),
                                         o, st.ctypes.data_as(ctypes.c_void_p), info, err, 256)
        if rc:
            raise RuntimeError(err.value.decode())
        self.last_info = {"n_passes": info[0], "n_gates": info[1], "n_kernel_gates": info[2], "n_mma_gates": info[3]}
        return st, info[0], info[1]

    def bitperm(self, psi, perm, opts=None):
        dt = 0 if psi.dtype == np.complex64 else 1
        n = int(round(np.log2(psi.size)))
        st = np.ascontiguousarray(psi).copy()
        p = np.ascontiguousarray(perm, dtype=np.uint32)
        o = _emu_opts(opts)
        info = (ctypes.c_int * 3)()
        err = ctypes.create_string_buffer(256)
        rc = self.lib.hq_emu_bitperm(dt, n, p.ctypes.data_as(ctypes.c_void_p), o,
                                     st.ctypes.data_as(ctypes.c_void_p), info, err, 256)
        if rc:
            raise RuntimeError(err.value.decode())
        return st, info[0]

    def bank_model(self, dtype, n, gates_pos, opts=None):
        """[(kind, ideal wavefronts, modelled wavefronts)] per kernel matrix (hq_emu_bank_model)."""
        ks = np.array([len(p) for p in gates_pos], dtype=np.uint32)
        pos = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint32) for p in gates_pos]))
        out = np.zeros(3 * (len(gates_pos) + 1), dtype=np.uint32)
        ng = self.lib.hq_emu_bank_model(dtype, n, len(gates_pos), ks.ctypes.data_as(ctypes.c_void_p),
                                        pos.ctypes.data_as(ctypes.c_void_p), _emu_opts(opts),
                                        out.ctypes.data_as(ctypes.c_void_p), out.size)
        if ng < 0:
            raise RuntimeError("bank model failed")
        return [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(ng)]

    def plan(self, dtype, n, gates_pos, opts=None):
        """Planner only: returns list of passes {tile_bits, n_high, n_gates, has_perm, high_pos, gate_ids}."""
        ks = np.array([len(p) for p in gates_pos], dtype=np.uint32)
        pos = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.uint32) for p in gates_pos]))
        o = _emu_opts(opts)
        out = np.zeros(64 * (len(gates_pos) + 4), dtype=np.uint32)
        w = self.lib.hq_emu_plan_dump(dtype, n, len(gates_pos), ks.ctypes.data_as(ctypes.c_void_p),
                                      pos.ctypes.data_as(ctypes.c_void_p), o,
                                      out.ctypes.data_as(ctypes.c_void_p), out.size)
        if w < 0:
            raise RuntimeError("plan failed")
        passes, i = [], 0
        while i < w:
            T, h, nk, hp, ng = (int(x) for x in out[i:i + 5])
            i += 5
            high = [int(x) for x in out[i:i + h]]
            i += h
            ids = [int(x) for x in out[i:i + ng]]
            i += ng
            passes.append(dict(tile_bits=T, n_high=h, n_gates=ng, n_kernel_gates=nk, has_perm=hp,
                               high_pos=high, gate_ids=ids))
        return passes
