"""CPU model of the CUDA kernel phases (hq_tile.cuh compiled by g++, driven by hq_emu.cpp) and
the planner, checked against the oracle.  No GPU involved; this only validates index math."""
import numpy as np
import pytest

from helpers import Emu, golden_gates, lower, initial_from, product_state, TOL


@pytest.fixture(scope="module")
def emu():
    return Emu()


def _rand_state(rng, n, ctype):
    psi = (rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)).astype(ctype)
    return (psi / np.linalg.norm(psi)).astype(ctype)


def _rand_gate(rng, n, k):
    U = (rng.standard_normal((2 ** k, 2 ** k)) + 1j * rng.standard_normal((2 ** k, 2 ** k))) / 2 ** (k / 2)
    return U, rng.permutation(n)[:k]


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [3, 7, 11, 13])
def test_random_circuits_vs_oracle(emu, oracle, c_oracle, ctype, n):
    rng = np.random.default_rng(100 + n)
    tol = 2e-5 if ctype == "complex64" else 1e-12
    for trial in range(6):
        psi = _rand_state(rng, n, ctype)
        gates = [_rand_gate(rng, n, int(rng.integers(1, min(n, 7) + 1))) for _ in range(int(rng.integers(1, 14)))]
        ref = oracle.evolve_oracle(psi, [(u.astype(ctype), p) for u, p in gates], c_oracle)
        for opts in (None, (8, 2, 1, 0, 0), (8, 3, 0, 0, 0), (9, 1, 1, 4, 0), (13, 5, 1, 0, 0),
                     (9, 2, 1, 0, 0, 0, -1), (10, 2, 1, 0, 0, 2, 0), (12, 3, 1, 0, 0, 3, 30),
                     (11, 2, 1, 0, 0, 2, -1, 0), (12, 3, 1, 0, 0, 3, -1, 1, 2), (13, 4, 1, 0, 0, 4, -1, 1, 2),
                     (9, 2, 1, 0, 0, 2, -1, 1, 0), (10, 2, 0, 0, 0, 0, -1, 1, 2)):
            out, n_pass, n_gates = emu.run(psi, gates, opts)
            assert n_gates == len(gates)
            assert np.abs(out - ref).max() < tol, (ctype, n, trial, opts)


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_every_k_every_low_bit(emu, oracle, c_oracle, ctype):
    """k = 1..8 with targets including bit 0 (inside the 16-byte unit for complex64)."""
    rng = np.random.default_rng(5)
    n = 12
    for k in range(1, 9):
        for low in (True, False):
            psi = _rand_state(rng, n, ctype)
            pool = np.arange(1, n)
            pos = list(rng.permutation(pool)[:k - 1]) + ([0] if low else [int(rng.permutation(pool)[-1])])
            pos = [int(x) for x in rng.permutation(pos)]
            if len(set(pos)) != k:
                continue
            U = _rand_gate(rng, n, k)[0]
            U2 = _rand_gate(rng, n, k)[0]
            ref = oracle.evolve_oracle(psi, [(U.astype(ctype), pos)], c_oracle)
            ref2 = oracle.evolve_oracle(ref, [(U2.astype(ctype), pos[::-1])], c_oracle)
            tol = 2e-5 if ctype == "complex64" else 1e-12
            for T in (10, 12):
                out, _, _ = emu.run(psi, [(U, pos)], (T, 1, 0, 0, 0, -1, -1, 1, 0))        # FMA paths
                assert np.abs(out - ref).max() < tol, (k, low, T)
                assert emu.last_info["n_mma_gates"] == 0
                # tensor-core path (k = 2..6): two unmerged gates in one pass (a lone k <= 2 gate keeps
                # its plain matrix for the direct kernel)
                out, n_pass, _ = emu.run(psi, [(U, pos), (U2, pos[::-1])], (T, 1, 1, 0, 0, 0, -1, 1, 2))
                assert np.abs(out - ref2).max() < tol, (k, low, T, "mma")
                assert n_pass == 1 and emu.last_info["n_mma_gates"] == (2 if 2 <= k <= 6 else 0), (k, low, T)


def test_golden_circuits_through_emu(emu, golden):
    z = golden["simulate"]
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"s{i}_ctype"])
        gq, n = golden_gates(z, f"s{i}")
        gates = lower(gq, n)
        init = initial_from(z, f"s{i}_init", n, ctype)
        psi0 = product_state(init, ctype) if isinstance(init, str) else init
        out, n_pass, _ = emu.run(psi0.astype(ctype), gates, (9, 3, 1, 0, 0))
        assert np.abs(out - z[f"s{i}_out"]).max() < 4 * TOL[ctype]
        assert n_pass < len(gates)                    # fusion happened


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_bitperm(emu, oracle, ctype):
    rng = np.random.default_rng(9)
    n = 13
    psi = _rand_state(rng, n, ctype)
    for trial in range(10):
        if trial % 2:
            perm = rng.permutation(n)                          # needs several passes
        else:
            m = int(rng.integers(2, 9))
            perm = np.concatenate([rng.permutation(m), np.arange(m, n)])
        ref = oracle.numpy_swap(psi, perm)                     # new bit i <- old bit perm[i]
        for opts in (None, (8, 2, 1, 0, 0), (10, 4, 1, 0, 0)):
            out, n_pass = emu.bitperm(psi, perm, opts)
            assert np.array_equal(out, ref), (trial, opts)     # pure data movement: bit-exact


def test_planner_properties(emu):
    """Every gate exactly once; gates sharing a bit keep their order; tiles fit."""
    rng = np.random.default_rng(11)
    for dtype, V in ((0, 1), (1, 0)):
        for n in (14, 20, 30):
            gates = [list(rng.permutation(n)[:int(rng.integers(1, 5))]) for _ in range(300)]
            for opts in (None, (12, 5, 1, 0, 0), (13, 4, 1, 8, 0), (11, 5, 0, 0, 0)):
                passes = emu.plan(dtype, n, gates, opts)
                order = [g for p in passes for g in p["gate_ids"]]
                assert sorted(order) == list(range(len(gates)))
                when = {g: i for i, g in enumerate(order)}
                for a in range(len(gates)):
                    for b in range(a + 1, len(gates)):
                        if set(gates[a]) & set(gates[b]):
                            assert when[a] < when[b]
                for p in passes:
                    L = p["tile_bits"] - p["n_high"]
                    assert L >= V and p["tile_bits"] <= min(n, 12 + V)
                    for g in p["gate_ids"]:
                        for bit in gates[g]:
                            assert bit < L or bit in p["high_pos"]
                if opts and opts[2] == 0:
                    assert all(p["n_gates"] == 1 for p in passes)


def test_scalar_plus_rank_one_gates(emu, oracle, c_oracle):
    """HQ_GATE_DR1: a super-operator of the form lambda * 1 + u v^T -- what the reference's dense
    depolarizing channel matrices are (hybridq/noise/channel/channel.py:413-529) -- is detected by the planner and
    applied with 2 * 2^k MACs per group; k = 3, 4, with and without a target on amplitude bit 0, both precisions."""
    import hybridq_b200 as hb
    rng = np.random.default_rng(9)
    paulis = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])]
    depol2 = sum(((1 - 0.01) if a == b == 0 else 0.01 / 15) * np.kron(np.kron(paulis[a], paulis[b]), np.kron(paulis[a], paulis[b]).conj())
                 for a in range(4) for b in range(4))
    u, v = rng.standard_normal(8) + 1j * rng.standard_normal(8), rng.standard_normal(8) + 1j * rng.standard_normal(8)
    generic3 = (0.9 + 0.1j) * np.eye(8) + 0.05 * np.outer(u, v)
    n = 12
    for ctype, tol in (("complex64", 1e-6), ("complex128", 1e-14)):
        for U in (depol2, generic3):
            k = int(np.log2(U.shape[0]))
            for pos in ([0, 3, 5, 9][:k], [2, 4, 7, 11][:k], [11, 1, 6, 3][:k]):
                psi = _rand_state(rng, n, ctype)
                other = (rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))) / 4
                gates = [(other, [1, 8]), (U, pos), (other.T, [8, 10])]
                ref = oracle.evolve_oracle(psi, [(g.astype(ctype), p) for g, p in gates], c_oracle)
                out, n_pass, _ = emu.run(psi, gates, None)
                assert np.abs(out - ref).max() < tol, (ctype, k, pos)
                plan = hb.Plan(gates, n, ctype)
                assert plan.arithmetic()[f"k{k}"].get("scalar_plus_rank_one") == 1, plan.arithmetic()
                # the channel (u = v = vectorised identity: 4 of 16 entries) runs in the sparse form, its scalar
                # carried by one of the dense gates; a dense u v^T does not
                assert plan.n_sparse_rank_one == (1 if U is depol2 else 0)
            assert hb.Plan([(U, [0, 3, 5, 9][:k])], n, ctype).n_sparse_rank_one == 0       # nothing to carry the scalar
    # a dense Haar matrix is not of that form
    from hybridq_b200.circuits import haar_unitary
    assert "scalar_plus_rank_one" not in str(hb.Plan([(haar_unitary(16, rng), [1, 3, 5, 7]), (other, [2, 9])], n, "complex64").arithmetic())
