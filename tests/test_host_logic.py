"""Host-side logic that needs neither the GPU nor the oracle."""
import numpy as np
import pytest

from hybridq_b200.circuits import (GateApply, matching_circuit, ksweep_circuit, sharded_circuit,
                                   to_positions)


def test_matching_circuit_is_seeded_and_well_formed():
    a = matching_circuit(20, depth=20, seed=20)
    b = matching_circuit(20, depth=20, seed=20)
    assert len(a) == len(b) and 250 < len(a) < 350
    for g, h in zip(a, b):
        assert g.qubits == h.qubits and np.array_equal(g.U, h.U)
        d = 2 ** len(g.qubits)
        assert np.allclose(g.U @ g.U.conj().T, np.eye(d), atol=1e-12)


def test_positions_follow_reference_convention():
    # first sorted qubit is the most significant bit; gate.qubits[0] is the matrix MSB
    g = GateApply(np.eye(4), ("a", "c"))
    h = GateApply(np.eye(2), ("b",))
    lowered, n = to_positions([g, h])
    assert n == 3
    assert lowered[0][1] == [0, 2]      # reversed(('a','c')) -> c is bit 0, a is bit 2
    assert lowered[1][1] == [1]


def test_sharded_circuit_fraction():
    gates = sharded_circuit(30, 3, depth=20, frac_global=0.2, seed=36)
    frac = np.mean([any(q < 3 for q in g.qubits) for g in gates])
    assert 0.1 < frac < 0.3


def test_ksweep():
    for k in range(1, 7):
        gs = ksweep_circuit(16, k, n_gates=5)
        assert all(len(set(g.qubits)) == k and g.U.shape == (2 ** k, 2 ** k) for g in gs)


def test_simulate_argument_errors_mirror_the_reference():
    import hybridq_b200 as hb
    gates = matching_circuit(12, depth=1, seed=1)
    with pytest.raises(ValueError, match="initial_state"):
        hb.simulate(gates, initial_state=None)
    with pytest.raises(ValueError, match="Wrong number of qubits"):
        hb.simulate(gates, initial_state="000")
    with pytest.raises(ValueError, match="not allowed"):
        hb.simulate(gates, initial_state="x" * 12)
    with pytest.raises(ValueError, match="dimension 2"):
        hb.simulate(gates, initial_state=np.zeros((3,) * 12))
    with pytest.raises(MemoryError):
        hb.simulate(gates, initial_state="0", max_largest_intermediate=2 ** 10)
    with pytest.raises(ValueError, match="tensor_only"):
        hb.simulate(gates, initial_state="0", tensor_only=True)
    with pytest.raises(NotImplementedError):
        hb.simulate(gates, initial_state="0", optimize="tn")


def test_lane_mapping_is_bank_conflict_free():
    """The planner's lane choices against a bank model of the shared-memory accesses (hq_tile.cuh swizzle with 7
    distinct bank vectors): FMA register-path gates are conflict-free for every target set; tensor-core gates
    are conflict-free unless their lane bits are forced onto colliding vectors (k = 2 with one of the 5
    colliding unit-bit pairs 0/7, 1/8, 2/9, 3/10, 4/11) -- and then cost at most 2 wavefronts per quarter-warp."""
    from helpers import Emu
    emu = Emu()
    rng = np.random.default_rng(2)
    collide = {(0, 7), (1, 8), (2, 9), (3, 10), (4, 11)}
    for dtype, V, T in ((0, 1, 13), (1, 0, 12)):
        n = T
        for k in (1, 2, 3, 4):
            for _ in range(40):
                pos = sorted(int(x) for x in rng.permutation(n)[:k])
                # two gates on the same bits, unmerged, so that k = 2 keeps the requested kind
                for mma in (0, 2):
                    if mma and k < 2:
                        continue
                    res = emu.bank_model(dtype, n, [pos, pos], (T, 1, 1, 0, 0, 0, -1, 0, mma))
                    for kind, ideal, actual in res:
                        if kind == 3:       # HQ_GATE_MMA
                            unit_bits = [p - V for p in pos if p - V >= 0]
                            forced = k == 2 and len(unit_bits) == 2 and tuple(unit_bits) in collide
                            if forced:
                                assert actual <= 2 * ideal, (dtype, pos, ideal, actual)
                            else:               # unit path and complex64 amplitude path (8-byte accesses)
                                assert actual == ideal, (dtype, pos, ideal, actual)
                        elif kind == 0:     # register path
                            assert actual == ideal, (dtype, pos, ideal, actual)


def test_merging_follows_the_measured_cost_table():
    """In-pass merging (hq_plan.cpp measured_cost).  complex128 (tensor-core path from k = 2): two k = 2 gates
    sharing a bit become one k = 3 matrix.  complex64 (FFMA2 slots up to k = 3, measured 0.62 / 1.36 ms): a chain of
    two k = 2 gates stays two matrices (1.24 < 1.36), a triangle of three becomes one k = 3 (the greedy pairwise
    merge is granted 10 % slack so that it can get there, and a second look splits the clusters that did not pay).
    Always: disjoint k = 2 gates stay apart, a 1-qubit gate is absorbed by a neighbour."""
    from helpers import Emu
    emu = Emu()
    n = 14
    for dtype in (0, 1):
        def mats(gates, opts=None):
            return sum(p["n_kernel_gates"] for p in emu.plan(dtype, n, gates, opts))
        assert mats([[3, 5], [5, 8]]) == (2 if dtype == 0 else 1)      # chain -> k = 3 only where that is cheaper
        assert mats([[3, 5], [7, 8]]) == 2                      # disjoint -> two matrices
        assert mats([[3, 5], [5]]) == 1 and mats([[4], [4, 9]]) == 1
        assert mats([[3, 5], [5, 8], [8, 3]]) == 1              # triangle on 3 bits -> one k = 3
        assert mats([[3, 5], [5, 8]], (0, -1, 1, 0, 0, -1, -1, 1, 0)) == 2      # FMA only: no k = 3 merge of a chain
        assert mats([[3, 5], [5, 8]], (0, -1, 1, 0, 0, 0, -1, 1, -1)) == 2      # merging off
        assert mats([[1, 2, 3], [2, 3, 4]]) == 1                # two k = 3 sharing two bits -> k = 4


def test_big_complex64_gates_get_a_pass_of_their_own():
    """A dense complex64 gate of 4..6 qubits opens a "solo" pass (one launch of the tcgen05 kernel on the GPU): only
    gates acting inside its qubits join, and they are multiplied into its matrix; everything else is fused around
    it as usual.  complex128 (no tcgen05 path) keeps fusing such gates into tile passes."""
    from helpers import Emu
    emu = Emu()
    n = 14
    gates = [[0, 1], [5, 9], [1, 4, 7, 9, 12], [4, 12], [7], [2, 3], [9, 13], [0, 1, 2, 3]]
    passes = emu.plan(0, n, gates, None)
    by_gate = {g: i for i, p in enumerate(passes) for g in p["gate_ids"]}
    assert sorted(by_gate) == list(range(len(gates)))
    solo5, solo4 = passes[by_gate[2]], passes[by_gate[7]]
    assert sorted(solo5["gate_ids"]) == [2, 3, 4] and solo5["n_kernel_gates"] == 1      # absorbed [4, 12] and [7]
    assert solo4["gate_ids"] == [7] and solo4["n_kernel_gates"] == 1
    assert by_gate[0] < by_gate[2] < by_gate[6] < by_gate[7]                             # order of overlapping gates kept
    assert len(emu.plan(1, n, gates, None)) == 1                                         # complex128: one fused pass
    assert len(emu.plan(0, 11, [[0, 1], [1, 4, 7, 9, 10]], None)) == 1                   # fewer than k + 7 qubits: fused


def test_plan_sharded_rejects_gates_wider_than_a_shard():
    """plan_sharded() used to loop forever when a gate touched more qubits than a rank holds."""
    import numpy as np
    import pytest
    from hybridq_b200.dist import plan_sharded
    with pytest.raises(ValueError):
        plan_sharded([(np.eye(8), [0, 1, 2])], n=5, g=3)
    with pytest.raises(ValueError):
        plan_sharded([(np.eye(8), [0, 1, 2])], n=4, g=2)
    with pytest.raises(ValueError):
        plan_sharded([(np.eye(2), [0])], n=2, g=2)
    ops, stats, where = plan_sharded([(np.eye(8), [0, 1, 4])], n=5, g=2)     # k = n - g: just fits
    assert stats["exchanges"] >= 1 and where == list(range(5))


def test_tcgen05_lane_map_exhaustive(tmp_path):
    """The memory-order lane map of the tcgen05 gate kernel (csrc/hq_umma.cuh) checked on the host for every target
    set of 4..6 bits among the 12 lowest amplitude bits: bijection onto the tile, conflict-free shared-memory
    quarter-warps with the skewed chunk stride, never more global lines than the row map
    (tools/umma_lane_map_check.cu: only the __host__ __device__ helpers run, nvcc builds it without a GPU)."""
    import json
    import shutil
    import subprocess
    from helpers import ROOT
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not shutil.which(nvcc):
        import pytest
        pytest.skip("nvcc not found")
    exe = tmp_path / "umma_lane_map_check"
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-o", str(exe),
                    str(ROOT / "tools" / "umma_lane_map_check.cu")], check=True, capture_output=True, timeout=300)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["failures"] == 0 and res["target_sets"] >= 2200 and res["with_chunk_lanes"] > 1000
    assert res["lines_per_quarter_memory_order_map"] < 0.6 * res["lines_per_quarter_rows_map"]


def test_tcgen05_operand_blocks_of_a_matrix():
    """umma_pack_matrix (hq_plan.cpp): the real form of U as K-major 16-byte units, split into TF32 hi / lo.  Checked
    against a numpy model of the layout: unit (n, c) at index c * R + n holds Bs[n][4c .. 4c + 3], Bs[2i + a][2j + b] =
    the real 2 x 2 block of U[i][j]; hi has 10 mantissa bits, hi + lo reproduces the fp32 entry to 2^-21."""
    import ctypes
    from helpers import Emu
    emu = Emu()
    rng = np.random.default_rng(8)
    for k in (4, 5, 6):
        dim, R = 2 ** k, 2 * 2 ** k
        U = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim))) / np.sqrt(dim)
        flat = np.ascontiguousarray(U.astype(np.complex128).reshape(-1)).view(np.float64)
        hi = np.zeros(R * R, np.float32)
        lo = np.zeros(R * R, np.float32)
        emu.lib.hq_emu_umma_pack(flat.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(k),
                                 hi.ctypes.data_as(ctypes.c_void_p), lo.ctypes.data_as(ctypes.c_void_p))
        U32 = U.astype(np.complex64)
        Bs = np.zeros((R, R), np.float32)
        Bs[0::2, 0::2], Bs[0::2, 1::2] = U32.real, -U32.imag        # re_out = re * re_in - im * im_in
        Bs[1::2, 0::2], Bs[1::2, 1::2] = U32.imag, U32.real         # im_out = im * re_in + re * im_in
        got_hi = hi.reshape(R // 4, R, 4).transpose(1, 0, 2).reshape(R, R)          # [c][n][e] -> [n][4c + e]
        got_lo = lo.reshape(R // 4, R, 4).transpose(1, 0, 2).reshape(R, R)
        assert np.all((got_hi.view(np.uint32) & 0x1fff) == 0) and np.all((got_lo.view(np.uint32) & 0x1fff) == 0)
        assert np.abs(got_hi - Bs).max() <= 2.0 ** -11 * np.abs(Bs).max()
        assert np.abs(got_hi.astype(np.float64) + got_lo - Bs).max() <= 2.0 ** -21 * np.abs(Bs).max()
