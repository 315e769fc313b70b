"""The oracle (oracle/hq_oracle.c + numpy restatement) is pinned against every golden vector
produced by the unmodified reference (tests/golden/make_golden.py) and, when oracle/_ref is
present, against the reference's own compiled core."""
import numpy as np
import pytest

from helpers import TOL, golden_gates, lower, initial_from, product_state, functional_items


def _tol(ctype, k=1):
    return {"complex64": 2e-6, "complex128": 1e-13}[str(np.dtype(ctype))]


def test_apply_u_golden(oracle, c_oracle, golden):
    z = golden["apply_u"]
    for i in range(int(z["n_cases"])):
        psi, U, pos, ref = z[f"c{i}_psi"], z[f"c{i}_U"], z[f"c{i}_pos"], z[f"c{i}_out"]
        planes = oracle.split_state(psi)
        assert c_oracle.apply_U(planes[0], planes[1], U, pos) == 0
        out = c_oracle.to_complex(planes[0], planes[1])
        assert np.abs(out - ref).max() < _tol(psi.dtype), (i, len(pos))
        out2 = oracle.numpy_apply_U(psi, U, pos)
        assert np.abs(out2 - ref).max() < _tol(psi.dtype), (i, len(pos))
    assert int(z["rc_pos_below_pack"]) == 1          # the reference's error contract


def test_swap_golden(oracle, c_oracle, golden):
    z = golden["swap"]
    n = int(z["n"])
    for i in range(int(z["n_cases"])):
        dt = str(z[f"c{i}_dtype"])
        a = (np.arange(2 ** n, dtype=np.int64) * 7 + 3).astype(dt)
        b = a.copy()
        assert c_oracle.swap(b, z[f"c{i}_pos"]) == 0
        assert np.array_equal(b, z[f"c{i}_out"]), i           # bit-exact
        assert np.array_equal(oracle.numpy_swap(a, z[f"c{i}_pos"]), z[f"c{i}_out"]), i


def test_simulate_golden(oracle, c_oracle, golden):
    z = golden["simulate"]
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"s{i}_ctype"])
        gq, n = golden_gates(z, f"s{i}")
        gates = lower(gq, n)
        init = initial_from(z, f"s{i}_init", n, ctype)
        psi0 = product_state(init, ctype) if isinstance(init, str) else init
        out = oracle.evolve_oracle(psi0, [(U.astype(ctype), p) for U, p in gates], c_oracle)
        # compress=4 cases merge matrices before applying them: same state up to rounding
        assert np.abs(out - z[f"s{i}_out"]).max() < 4 * TOL[ctype], (i, str(z[f"s{i}_tag"]))


def test_expectation_golden(oracle, c_oracle, golden):
    """<state| op |state> of the reference's expectation_value (simulation.py:1125-1217)."""
    z = golden["expectation"]
    n = int(z["n_qubits"])
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"e{i}_ctype"])
        gates = [(z[f"e{i}_g{j}_U"], z[f"e{i}_g{j}_q"]) for j in range(int(z[f"e{i}_ngates"]))]
        state = z[f"e{i}_state"].astype(ctype)
        out = oracle.evolve_oracle(state, [(U.astype(ctype), p) for U, p in lower(gates, n)], c_oracle)
        val = complex(np.real_if_close(np.vdot(state, out)))     # as the reference returns it (:1217)
        assert abs(val - complex(z[f"e{i}_value"])) <= (2e-6 if ctype == "complex64" else 1e-12)


def test_functional_golden(oracle, c_oracle, golden):
    """Projection / Measure restatements against the reference's simulate() with FunctionalGates."""
    z = golden["functional"]
    n = int(z["n_qubits"])
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"f{i}_ctype"])
        psi = product_state("+" * n, ctype)
        np.random.seed(int(z[f"f{i}_seed"]))
        for kind, payload, q in functional_items(z, i):
            if kind == "U":
                psi = oracle.evolve_oracle(psi, [(payload.astype(ctype), [n - 1 - x for x in reversed(q)])], c_oracle)
            elif kind == "P":
                psi = oracle.numpy_project(psi, [n - 1 - x for x in q], [int(c) for c in payload])
            else:
                psi, _ = oracle.numpy_measure(psi, [n - 1 - x for x in q])
        assert np.abs(psi - z[f"f{i}_out"]).max() < 4 * TOL[ctype], i


def test_dot_golden(oracle, golden):
    z = golden["dot"]
    n = int(z["n"])
    for i in range(int(z["n_cases"])):
        psi, U, axes, ref = z[f"d{i}_psi"], z[f"d{i}_U"], z[f"d{i}_axes"], z[f"d{i}_out"]
        pos = (n - axes[::-1] - 1)                       # dot.py:215
        out = oracle.numpy_apply_U(psi, U, pos)
        assert np.abs(out - ref).max() < _tol(psi.dtype), i


def test_transpose_golden(oracle, golden):
    z = golden["transpose"]
    n = int(z["n"])
    for i in range(int(z["n_cases"])):
        dt = str(z[f"t{i}_dtype"])
        a = (np.arange(2 ** n, dtype=np.int64) * 5 + 1).astype(dt).reshape((2,) * n)
        axes = z[f"t{i}_axes"]
        assert np.array_equal(np.transpose(a, axes).reshape(-1), z[f"t{i}_out"])
        n_ord = next(j for j, x in enumerate(axes) if j != x)
        pos = n - axes[n_ord:][::-1] - 1                 # transpose.py:139
        assert np.array_equal(oracle.numpy_swap(a.reshape(-1), pos), z[f"t{i}_out"])


def test_dm_golden(oracle, c_oracle, golden):
    z = golden["dm"]
    n = int(z["n_super"])
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"m{i}_ctype"])
        ng = int(z[f"m{i}_ngates"])
        gates = [(z[f"m{i}_g{j}_U"].astype(ctype), [n - 1 - int(x) for x in reversed(z[f"m{i}_g{j}_q"])])
                 for j in range(ng)]
        init = z[f"m{i}_init"]
        psi0 = product_state(str(init), ctype) if init.dtype.kind in "US" else init.astype(ctype)
        out = oracle.evolve_oracle(psi0, gates, c_oracle)
        assert np.abs(out - z[f"m{i}_out"]).max() < TOL[ctype]
        rho = out.reshape(2 ** (n // 2), 2 ** (n // 2))
        assert abs(np.trace(rho) - 1) < 1e-5


@pytest.mark.parametrize("variant", ["avx2", "wheel"])
def test_oracle_vs_compiled_reference(oracle, c_oracle, variant):
    if not oracle.RefCore.available(variant):
        pytest.skip(f"oracle/_ref/{variant} not built")
    core = oracle.RefCore(variant)
    rng = np.random.default_rng(7)
    n = 13
    for ctype in ("complex64", "complex128"):
        psi = (rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)).astype(ctype)
        psi /= np.linalg.norm(psi)
        gates = []
        for _ in range(25):
            k = int(rng.integers(1, 6))
            U = (rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))) / 2 ** (k / 2)
            gates.append((U.astype(ctype), rng.permutation(n)[:k]))
        ref = oracle.evolve_ref(psi, gates, core)            # includes the low-bit swap bookkeeping
        out = oracle.evolve_oracle(psi, gates, c_oracle)
        # non-unitary gates: compare relative to the largest amplitude
        assert np.abs(out - ref).max() / np.abs(ref).max() < 10 * _tol(ctype)


def test_to_complex(oracle, c_oracle):
    rng = np.random.default_rng(3)
    for ft in (np.float32, np.float64):
        re = rng.standard_normal(1000).astype(ft)
        im = rng.standard_normal(1000).astype(ft)
        assert np.array_equal(c_oracle.to_complex(re, im), re + 1j * im)
