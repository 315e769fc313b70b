#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

What produces each file (all outputs come from reference code, none from this repo's
oracle or CUDA path):

  apply_u.npz    raw ctypes calls of the reference core `apply_U_float32/64`
                 (/root/reference/include/python_U.cpp:131-143) compiled into
                 oracle/_ref/avx2 -- k = 1..6, complex64 and complex128.
  swap.npz       raw ctypes calls of `swap_*` (/root/reference/include/python_swap.cpp:70-98).
  simulate.npz   `hybridq.circuit.simulation.simulate(..., optimize='evolution')`
                 (/root/reference/hybridq/circuit/simulation/simulation.py:59) on seeded
                 circuits: compress=0 and the default compress=4, both precisions,
                 string and array initial states, tuple qubit labels.
  dot.npz        `hybridq.utils.dot` through the C++ core (dot.py:139, raise_if_hcore_fails).
  transpose.npz  `hybridq.utils.transpose` through the C++ core (transpose.py:61).
  functional.npz   `simulate(optimize='evolution')` on circuits with Projection and Measure FunctionalGates
                 (simulation.py:525-554, gate/projection.py, gate/measure.py), numpy's generator seeded.
  expectation.npz  `hybridq.circuit.simulation.expectation_value` (simulation.py:1125) on 12-qubit states.
  dm_large.npz   the same at 10 and 12 qubits (2^20 / 2^24 superkets): lowered gates + sampled amplitudes.
  dm.npz         `hybridq.dm.circuit.simulation.simulate` (dm/circuit/simulation.py:118) on a
                 6-qubit circuit with depolarizing noise; the lowered 12-"qubit" circuit that
                 it hands to `simulate` is captured and stored as (matrix, qubit-index) lists.

The reference's `simulation` module imports opt_einsum and more_itertools at module
level (simulation.py:41-43); neither is installed here nor used by the evolution path,
so two import stubs from oracle/_ref/stubs are put on PYTHONPATH.  This is declared in
DESIGN.md.  The script re-executes itself with LD_LIBRARY_PATH pointing at the
reference build, because the reference resolves 'hybridq.so' by bare name at import.
"""
import os
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
REFLIB = ROOT / "oracle" / "_ref" / "avx2"
STUBS = ROOT / "oracle" / "_ref" / "stubs"

if os.environ.get("HQ_GOLDEN_CHILD") != "1":
    if not REF.exists() or not (REFLIB / "hybridq.so").exists():
        sys.exit("need /root/reference and `make -C oracle ref` first")
    env = dict(os.environ)
    env["HQ_GOLDEN_CHILD"] = "1"
    env["LD_LIBRARY_PATH"] = f"{REFLIB}:" + env.get("LD_LIBRARY_PATH", "")
    env["PYTHONPATH"] = f"{REF}:{STUBS}:{ROOT}:" + env.get("PYTHONPATH", "")
    env["OMP_NUM_THREADS"] = "4"
    os.execve(sys.executable, [sys.executable, "-W", "ignore", __file__] + sys.argv[1:], env)

import numpy as np  # noqa: E402

from oracle import oracle as O  # noqa: E402  (only for RefCore ctypes binding + aligned buffers)
from hybridq_b200.circuits import matching_circuit, ksweep_circuit, haar_unitary  # noqa: E402

core = O.RefCore("avx2")
L = core.log2_pack_size
print("reference core log2_pack_size =", L)


def rand_state(rng, n, ctype):
    ft = np.float32 if ctype == "complex64" else np.float64
    psi = rng.standard_normal(2 ** n).astype(ft) + 1j * rng.standard_normal(2 ** n).astype(ft)
    return (psi / np.linalg.norm(psi)).astype(ctype)


# ---------------------------------------------------------------- apply_u.npz
def make_apply_u():
    rng = np.random.default_rng(1001)
    out = {}
    idx = 0
    n = 11
    for ctype in ("complex64", "complex128"):
        for k in range(1, 7):
            for variant in range(2):
                psi = rand_state(rng, n, ctype)
                # non-unitary dense matrix, as in the reference's own test_utils__dot
                U = (rng.standard_normal((2 ** k, 2 ** k)) +
                     1j * rng.standard_normal((2 ** k, 2 ** k))).astype(ctype) / 2 ** (k / 2)
                if variant == 0:
                    pos = rng.permutation(np.arange(L, n))[:k]
                else:  # the k highest bits, shuffled
                    pos = rng.permutation(np.arange(n - k, n))
                planes = O.split_state(psi)
                rc = core.apply_U(planes[0], planes[1], U, pos)
                assert rc == 0
                out[f"c{idx}_psi"] = psi
                out[f"c{idx}_U"] = U
                out[f"c{idx}_pos"] = pos.astype(np.uint32)
                out[f"c{idx}_out"] = core.to_complex(planes[0], planes[1])
                idx += 1
    # error contract: a position below the pack width is rejected with rc=1 (U.h:48-54)
    psi = rand_state(rng, n, "complex64")
    planes = O.split_state(psi)
    rc = core.apply_U(planes[0], planes[1], np.eye(2, dtype="complex64"), [0])
    out["rc_pos_below_pack"] = np.int32(rc)
    out["ref_log2_pack_size"] = np.int32(L)
    out["n_cases"] = np.int32(idx)
    np.savez_compressed(HERE / "apply_u.npz", **out)
    print("apply_u.npz:", idx, "cases; rc(pos<L) =", rc)


# ---------------------------------------------------------------- swap.npz
def make_swap():
    rng = np.random.default_rng(1002)
    out = {}
    idx = 0
    n = 12
    for dt in ("float32", "float64", "int32", "int64", "uint32", "uint64"):
        for m in (1, 2, 5, 8, 10, 12):
            pos = rng.permutation(m).astype(np.uint32)
            a = O.aligned_empty((2 ** n,), dt, 4096)
            a[:] = (np.arange(2 ** n, dtype=np.int64) * 7 + 3).astype(dt)
            rc = core.swap(a, pos)
            assert rc == 0
            out[f"c{idx}_dtype"] = np.array(dt)
            out[f"c{idx}_pos"] = pos
            out[f"c{idx}_out"] = a.copy()
            idx += 1
    out["n"] = np.int32(n)
    out["n_cases"] = np.int32(idx)
    np.savez_compressed(HERE / "swap.npz", **out)
    print("swap.npz:", idx, "cases")


# ---------------------------------------------------------------- simulate.npz
def make_simulate():
    from hybridq.gate import MatrixGate
    from hybridq.circuit import Circuit
    from hybridq.circuit.simulation import simulate
    out = {}
    idx = 0

    def store(tag, gates, qubits, init, psi, kw):
        nonlocal idx
        out[f"s{idx}_tag"] = np.array(tag)
        out[f"s{idx}_ngates"] = np.int32(len(gates))
        for j, g in enumerate(gates):
            out[f"s{idx}_g{j}_U"] = np.asarray(g.matrix())
            out[f"s{idx}_g{j}_q"] = np.array([qubits.index(q) for q in g.qubits], dtype=np.int32)
        out[f"s{idx}_nqubits"] = np.int32(len(qubits))
        out[f"s{idx}_init"] = np.array(init) if isinstance(init, str) else np.asarray(init)
        out[f"s{idx}_out"] = np.asarray(psi).reshape(-1)
        out[f"s{idx}_ctype"] = np.array(kw["complex_type"])
        idx += 1

    for n, depth, seed in ((12, 8, 12), (14, 6, 14)):
        gates = matching_circuit(n, depth=depth, seed=seed)
        qubits = list(range(n))
        for ctype in ("complex64", "complex128"):
            for compress in (0, 4):
                circ = Circuit(MatrixGate(g.U, qubits=list(g.qubits)) for g in gates)
                kw = dict(optimize="evolution", simplify=False, compress=compress,
                          complex_type=ctype)
                psi = simulate(circ, initial_state="0" * n, **kw)
                store(f"matching n={n} depth={depth} compress={compress}", gates, qubits,
                      "0" * n, psi, kw)
    # random initial state array + '+' string + k up to 3, tuple labels
    rng = np.random.default_rng(77)
    n = 12
    labels = [(i // 4, i % 4) for i in range(n)]          # sorted tuples
    gl = []
    for _ in range(30):
        k = int(rng.integers(1, 4))
        qs = [labels[int(i)] for i in rng.permutation(n)[:k]]
        gl.append(MatrixGate(haar_unitary(2 ** k, rng), qubits=qs))
    for ctype in ("complex64", "complex128"):
        init = rand_state(rng, n, ctype).reshape((2,) * n)
        kw = dict(optimize="evolution", simplify=False, compress=0, complex_type=ctype)
        psi = simulate(Circuit(gl), initial_state=init, **kw)
        store("tuple labels, array initial state", gl, labels, init.reshape(-1), psi, kw)
        psi = simulate(Circuit(gl), initial_state="+-01" * 3, **kw)
        store("tuple labels, '+-01' initial state", gl, labels, "+-01" * 3, psi, kw)
    out["n_cases"] = np.int32(idx)
    np.savez_compressed(HERE / "simulate.npz", **out)
    print("simulate.npz:", idx, "cases")


# ---------------------------------------------------------------- dot.npz / transpose.npz
def make_dot_transpose():
    from hybridq.utils import dot, transpose
    rng = np.random.default_rng(1003)
    out = {}
    idx = 0
    n = 12
    for ctype in ("complex64", "complex128"):
        for k in (2, 3, 4, 5, 6):
            psi = rand_state(rng, n, ctype).reshape((2,) * n)
            U = (rng.standard_normal((2 ** k, 2 ** k)) +
                 1j * rng.standard_normal((2 ** k, 2 ** k))).astype(ctype) / 2 ** (k / 2)
            axes = rng.permutation(n)[:k]
            res = dot(U, psi, axes_b=axes, raise_if_hcore_fails=True)
            out[f"d{idx}_psi"] = psi.reshape(-1)
            out[f"d{idx}_U"] = U
            out[f"d{idx}_axes"] = axes.astype(np.int32)
            out[f"d{idx}_out"] = np.asarray(res).reshape(-1)
            idx += 1
    out["n_cases"] = np.int32(idx)
    out["n"] = np.int32(n)
    np.savez_compressed(HERE / "dot.npz", **out)
    print("dot.npz:", idx, "cases")

    out = {}
    idx = 0
    for dt in ("float32", "float64", "int32", "int64", "uint32", "uint64"):
        for tail in (6, 12):
            a = (np.arange(2 ** n, dtype=np.int64) * 5 + 1).astype(dt).reshape((2,) * n)
            axes = np.concatenate([np.arange(n - tail), n - tail + rng.permutation(tail)])
            res = transpose(a, axes, raise_if_hcore_fails=True)
            assert np.array_equal(res, np.transpose(a, axes))
            out[f"t{idx}_dtype"] = np.array(dt)
            out[f"t{idx}_axes"] = axes.astype(np.int32)
            out[f"t{idx}_out"] = np.ascontiguousarray(res).reshape(-1)
            idx += 1
    out["n_cases"] = np.int32(idx)
    out["n"] = np.int32(n)
    np.savez_compressed(HERE / "transpose.npz", **out)
    print("transpose.npz:", idx, "cases")


# ---------------------------------------------------------------- dm.npz
def make_dm():
    import hybridq.dm.circuit.simulation as dmsim
    import hybridq.circuit.simulation as csim
    from hybridq.gate import MatrixGate
    from hybridq.circuit import Circuit
    from hybridq.noise.utils import add_depolarizing_noise

    captured = {}
    real_simulate = csim.simulate

    def spy(circuit, initial_state, **kw):
        captured["circuit"] = list(circuit)
        captured["initial_state"] = initial_state
        return real_simulate(circuit=circuit, initial_state=initial_state, **kw)

    nq = 6
    gates = matching_circuit(nq, depth=4, seed=606)
    circ = Circuit(MatrixGate(g.U, qubits=list(g.qubits)) for g in gates)
    noisy = add_depolarizing_noise(circ, probs=(0.001, 0.01))
    out = {}
    for ci, ctype in enumerate(("complex64", "complex128")):
        # dm.simulate does `from hybridq.circuit.simulation import simulate` inside the
        # function (dm/circuit/simulation.py:193), so patch the source module attribute.
        csim.simulate = spy
        try:
            rho = dmsim.simulate(noisy, initial_state="0", optimize="evolution",
                                 simplify=False, compress=0, complex_type=ctype)
        finally:
            csim.simulate = real_simulate
        from hybridq.circuit import utils as cutils
        lowered = list(cutils.flatten(Circuit(captured["circuit"])))   # TupleGate -> members
        qubits = sorted({q for g in lowered for q in g.qubits})
        assert len(qubits) == 2 * nq
        out[f"m{ci}_ngates"] = np.int32(len(lowered))
        for j, g in enumerate(lowered):
            out[f"m{ci}_g{j}_U"] = np.asarray(g.matrix())
            out[f"m{ci}_g{j}_q"] = np.array([qubits.index(q) for q in g.qubits], dtype=np.int32)
        init = captured["initial_state"]
        out[f"m{ci}_init"] = np.array(init) if isinstance(init, str) else np.asarray(init).reshape(-1)
        out[f"m{ci}_out"] = np.asarray(rho).reshape(-1)
        out[f"m{ci}_ctype"] = np.array(ctype)
        r = np.asarray(rho).reshape(2 ** nq, 2 ** nq)
        print(f"dm {ctype}: {len(lowered)} lowered gates, trace = {np.trace(r):.6f}, "
              f"k-hist = {np.bincount([len(g.qubits) for g in lowered])}")
    out["n_cases"] = np.int32(2)
    out["n_super"] = np.int32(2 * nq)
    np.savez_compressed(HERE / "dm.npz", **out)


# ---------------------------------------------------------------- dm15_circuit.npz (config 5)
def make_dm15():
    """BASELINE config 5: 15-qubit depth-10 circuit + depolarizing noise, lowered by the reference's
    dm front-end to a 30-"qubit" circuit.  Only the lowered gate list is stored (the 2^30 superket
    has no CPU oracle); the GPU run checks trace / hermiticity and is timed by tools/run_configs.py."""
    import hybridq.dm.circuit.simulation as dmsim
    import hybridq.circuit.simulation as csim
    from hybridq.gate import MatrixGate
    from hybridq.circuit import Circuit, utils as cutils
    from hybridq.noise.utils import add_depolarizing_noise

    class _Captured(Exception):
        pass

    captured = {}

    def spy(circuit, initial_state, **kw):
        captured["circuit"] = list(circuit)
        raise _Captured()

    nq = 15
    gates = matching_circuit(nq, depth=10, seed=1515)
    circ = Circuit(MatrixGate(g.U, qubits=list(g.qubits)) for g in gates)
    noisy = add_depolarizing_noise(circ, probs=(0.001, 0.01))
    real = csim.simulate
    csim.simulate = spy
    try:
        dmsim.simulate(noisy, initial_state="0", optimize="evolution", simplify=False, compress=0,
                       complex_type="complex64", max_largest_intermediate=2 ** 30)
    except _Captured:
        pass
    finally:
        csim.simulate = real
    lowered = list(cutils.flatten(Circuit(captured["circuit"])))
    qubits = sorted({q for g in lowered for q in g.qubits})
    assert len(qubits) == 2 * nq
    out = {"n_super": np.int32(2 * nq), "ngates": np.int32(len(lowered))}
    for j, g in enumerate(lowered):
        out[f"g{j}_U"] = np.asarray(g.matrix()).astype(np.complex64)
        out[f"g{j}_q"] = np.array([qubits.index(q) for q in g.qubits], dtype=np.int32)
    np.savez_compressed(HERE / "dm15_circuit.npz", **out)
    print(f"dm15_circuit.npz: {len(lowered)} lowered gates, k-hist = "
          f"{np.bincount([len(g.qubits) for g in lowered])}")


def make_functional():
    """FunctionalGates inside `simulate(optimize='evolution')` (simulation.py:525-554): Projection
    (gate/projection.py) and Measure (gate/measure.py; the draw uses numpy's global generator, seeded here)."""
    from hybridq.gate import MatrixGate, Projection, Measure
    from hybridq.circuit import Circuit
    from hybridq.circuit.simulation import simulate
    rng = np.random.default_rng(55)
    out = {}
    idx = 0
    n = 12
    for ctype in ("complex64", "complex128"):
        for seed in (123, 7):
            items = []
            def rand_gates(m):
                for _ in range(m):
                    k = int(rng.integers(1, 4))
                    items.append(("U", haar_unitary(2 ** k, rng), [int(i) for i in rng.permutation(n)[:k]]))
            rand_gates(10)
            items.append(("P", "10", [3, 7]))
            rand_gates(10)
            items.append(("M", None, [1, 5, 9]))
            rand_gates(6)
            items.append(("M", None, [0]))
            rand_gates(4)
            circ = Circuit(MatrixGate(U, qubits=q) if kind == "U" else
                           (Projection(state=U, qubits=q) if kind == "P" else Measure(qubits=q))
                           for kind, U, q in items)
            np.random.seed(seed)
            psi = simulate(circ, initial_state="+" * n, optimize="evolution", simplify=False, compress=0,
                           complex_type=ctype)
            out[f"f{idx}_ctype"] = np.array(ctype)
            out[f"f{idx}_seed"] = np.int32(seed)
            out[f"f{idx}_nitems"] = np.int32(len(items))
            for j, (kind, U, q) in enumerate(items):
                out[f"f{idx}_i{j}_kind"] = np.array(kind)
                out[f"f{idx}_i{j}_q"] = np.array(q, dtype=np.int32)
                if kind == "U":
                    out[f"f{idx}_i{j}_U"] = np.asarray(U)
                elif kind == "P":
                    out[f"f{idx}_i{j}_state"] = np.array(U)
            out[f"f{idx}_out"] = np.asarray(psi).reshape(-1)
            idx += 1
    out["n_cases"] = np.int32(idx)
    out["n_qubits"] = np.int32(n)
    np.savez_compressed(HERE / "functional.npz", **out)
    print("functional.npz:", idx, "cases; norms", [float(np.linalg.norm(out[f"f{i}_out"])) for i in range(idx)])


def make_expectation():
    """`hybridq.circuit.simulation.expectation_value` (simulation.py:1125-1217): <state| op |state> with
    op acting on a subset of the qubits (the reference pads it with identity gates)."""
    from hybridq.gate import MatrixGate
    from hybridq.circuit import Circuit
    from hybridq.circuit.simulation import expectation_value
    rng = np.random.default_rng(91)
    out = {}
    idx = 0
    n = 12
    qubits = list(range(n))
    for ctype in ("complex64", "complex128"):
        for n_gates, kmax in ((3, 2), (12, 3)):
            state = rand_state(rng, n, ctype).reshape((2,) * n)
            gl = []
            for _ in range(n_gates):
                k = int(rng.integers(1, kmax + 1))
                qs = [int(i) for i in rng.permutation(n - 2)[:k]]          # qubits n-2, n-1 stay untouched
                gl.append(MatrixGate(haar_unitary(2 ** k, rng), qubits=qs))
            val = expectation_value(state=state, op=Circuit(gl), qubits_order=qubits, complex_type=ctype,
                                    simplify=False, compress=0)
            out[f"e{idx}_ctype"] = np.array(ctype)
            out[f"e{idx}_state"] = state.reshape(-1)
            out[f"e{idx}_ngates"] = np.int32(len(gl))
            for j, g in enumerate(gl):
                out[f"e{idx}_g{j}_U"] = np.asarray(g.matrix())
                out[f"e{idx}_g{j}_q"] = np.array(g.qubits, dtype=np.int32)
            out[f"e{idx}_value"] = np.complex128(val)
            idx += 1
    out["n_cases"] = np.int32(idx)
    out["n_qubits"] = np.int32(n)
    np.savez_compressed(HERE / "expectation.npz", **out)
    print("expectation.npz:", idx, "cases", [complex(out[f"e{i}_value"]) for i in range(idx)])

# ---------------------------------------------------------------- dm_large.npz (config 5 parity at 2^20 / 2^24)
def make_dm_large():
    """SURVEY 8(d) config 5 parity: 10- and 12-qubit density matrices (2^20 / 2^24 superkets, complex64) through the
    reference's dm front-end and its evolution core.  The full output is too large to commit, so the file holds
    the lowered gate list (what dm.simulate hands to `simulate`), 8192 sampled amplitudes of the reference result,
    its trace and its squared norm; the GPU test recomputes the full vector with the reference core
    (oracle/_ref) on the box and additionally checks these samples."""
    import hybridq.dm.circuit.simulation as dmsim
    import hybridq.circuit.simulation as csim
    from hybridq.gate import MatrixGate
    from hybridq.circuit import Circuit, utils as cutils
    from hybridq.noise.utils import add_depolarizing_noise

    captured = {}
    real_simulate = csim.simulate

    def spy(circuit, initial_state, **kw):
        captured["circuit"] = list(circuit)
        captured["initial_state"] = initial_state
        return real_simulate(circuit=circuit, initial_state=initial_state, **kw)

    out = {"n_cases": np.int32(2)}
    for ci, nq in enumerate((10, 12)):
        gates = matching_circuit(nq, depth=6, seed=5000 + nq)
        circ = Circuit(MatrixGate(g.U, qubits=list(g.qubits)) for g in gates)
        noisy = add_depolarizing_noise(circ, probs=(0.001, 0.01))
        csim.simulate = spy
        try:
            rho = dmsim.simulate(noisy, initial_state="0", optimize="evolution", simplify=False, compress=0,
                                 complex_type="complex64")
        finally:
            csim.simulate = real_simulate
        lowered = list(cutils.flatten(Circuit(captured["circuit"])))
        qubits = sorted({q for g in lowered for q in g.qubits})
        assert len(qubits) == 2 * nq
        flat = np.asarray(rho).reshape(-1)
        rng = np.random.default_rng(77 + nq)
        idx = np.unique(np.concatenate([rng.integers(0, flat.size, 8192),
                                        np.arange(2 ** nq, dtype=np.int64) * (2 ** nq + 1)]))   # incl. the diagonal
        out[f"L{ci}_nq"] = np.int32(nq)
        out[f"L{ci}_ngates"] = np.int32(len(lowered))
        for j, g in enumerate(lowered):
            out[f"L{ci}_g{j}_U"] = np.asarray(g.matrix()).astype(np.complex64)
            out[f"L{ci}_g{j}_q"] = np.array([qubits.index(q) for q in g.qubits], dtype=np.int32)
        out[f"L{ci}_init"] = np.array(captured["initial_state"])
        out[f"L{ci}_idx"] = idx.astype(np.int64)
        out[f"L{ci}_val"] = flat[idx]
        out[f"L{ci}_trace"] = np.complex128(np.trace(np.asarray(rho).reshape(2 ** nq, 2 ** nq)))
        out[f"L{ci}_norm2"] = np.float64(np.vdot(flat.astype(np.complex128), flat.astype(np.complex128)).real)
        print(f"dm_large nq={nq}: {len(lowered)} lowered gates, trace = {out[f'L{ci}_trace']:.6f}, "
              f"k-hist = {np.bincount([len(g.qubits) for g in lowered])}")
    np.savez_compressed(HERE / "dm_large.npz", **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dm_large":
        make_dm_large()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dm15":
        make_dm15()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "expectation":
        make_expectation()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "functional":
        make_functional()
        sys.exit(0)
    make_apply_u()
    make_swap()
    make_simulate()
    make_dot_transpose()
    make_dm()
    make_dm15()
    make_dm_large()
    make_expectation()
    make_functional()
    for f in sorted(HERE.glob("*.npz")):
        print(f"{f.name}: {f.stat().st_size / 1e6:.2f} MB")
