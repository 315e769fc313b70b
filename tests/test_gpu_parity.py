"""GPU parity tests proper: everything goes through the C ABI of libhybridq_b200.so and is
compared with the oracle / the golden vectors of the unmodified reference.
Tolerances (BASELINE.json north_star): max-abs 1e-6 complex64, 1e-12 complex128."""
import ctypes

import numpy as np
import pytest

from helpers import TOL, golden_gates, lower, initial_from, product_state, functional_items

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hybridq_b200
    return hybridq_b200


def _rand_state(rng, n, ctype):
    psi = (rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)).astype(ctype)
    return (psi / np.linalg.norm(psi)).astype(ctype)


def _haar(rng, k):
    from hybridq_b200.circuits import haar_unitary
    return haar_unitary(2 ** k, rng)


# ------------------------------------------------------------------ Part 1: host-pointer ABI
def test_host_abi_apply_u_golden(hb, oracle, golden):
    z = golden["apply_u"]
    lib = hb.lib
    assert lib.get_log2_pack_size() == 1
    for i in range(int(z["n_cases"])):
        psi, U, pos, ref = z[f"c{i}_psi"], z[f"c{i}_U"], z[f"c{i}_pos"], z[f"c{i}_out"]
        ctype = str(psi.dtype)
        planes = oracle.split_state(psi, alignment=32)
        ct = ctypes.c_float if ctype == "complex64" else ctypes.c_double
        fn = lib.apply_U_float32 if ctype == "complex64" else lib.apply_U_float64
        p = ctypes.POINTER(ct)
        Uc = np.ascontiguousarray(U)
        posc = np.ascontiguousarray(pos, dtype=np.uint32)
        n = int(np.log2(psi.size))
        rc = fn(planes[0].ctypes.data_as(p), planes[1].ctypes.data_as(p), Uc.ctypes.data_as(p),
                posc.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), n, len(posc))
        assert rc == 0, hb._lib.last_error()
        out = planes[0] + 1j * planes[1]
        assert np.abs(out - ref).max() <= TOL[ctype], (i, len(pos))


def test_host_abi_swap_and_to_complex_golden(hb, oracle, golden):
    z = golden["swap"]
    n = int(z["n"])
    core = oracle.RefCore(path=hb.DROPIN_DIR)         # this repo's library under the reference's names
    for i in range(int(z["n_cases"])):
        dt = str(z[f"c{i}_dtype"])
        a = (np.arange(2 ** n, dtype=np.int64) * 7 + 3).astype(dt)
        assert core.swap(a, z[f"c{i}_pos"]) == 0
        assert np.array_equal(a, z[f"c{i}_out"]), i       # bit-exact
    rng = np.random.default_rng(0)
    for ft in (np.float32, np.float64):
        re = rng.standard_normal(5000).astype(ft)
        im = rng.standard_normal(5000).astype(ft)
        assert np.array_equal(core.to_complex(re, im), re + 1j * im)
        assert np.array_equal(hb.to_complex(re, im), re + 1j * im)


def test_reference_style_host_loop_over_the_dropin_library(hb, oracle, c_oracle, golden):
    """Drive swap_* / apply_U_* / to_complex* in the order the reference's own host loop
    would (simulation.py:522-675), on the golden circuits, with this library bound under the
    reference's file names."""
    core = oracle.RefCore(path=hb.DROPIN_DIR)
    assert core.log2_pack_size == 1
    z = golden["simulate"]
    for i in (0, 2, 8, 9):
        ctype = str(z[f"s{i}_ctype"])
        gq, n = golden_gates(z, f"s{i}")
        gates = [(U.astype(ctype), p) for U, p in lower(gq, n)]
        init = initial_from(z, f"s{i}_init", n, ctype)
        psi0 = product_state(init, ctype) if isinstance(init, str) else init
        out = oracle.evolve_ref(psi0, gates, core)
        assert np.abs(out - z[f"s{i}_out"]).max() <= 4 * TOL[ctype], i


# ------------------------------------------------------------------ Part 2: device-resident ABI
@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_device_apply_every_k_any_bits(hb, oracle, c_oracle, ctype):
    rng = np.random.default_rng(21)
    n = 15
    for k in range(1, 9):
        for variant in range(4):
            psi = _rand_state(rng, n, ctype)
            if variant == 0:
                pos = list(range(k))                                   # lowest bits incl. bit 0
            elif variant == 1:
                pos = list(range(n - k, n))                            # highest bits
            else:
                pos = [int(x) for x in rng.permutation(n)[:k]]
            pos = [int(x) for x in rng.permutation(pos)]
            U = _haar(rng, k).astype(ctype)
            planes = oracle.split_state(psi)
            assert c_oracle.apply_U(planes[0], planes[1], U, pos) == 0
            ref = c_oracle.to_complex(planes[0], planes[1])
            for use_direct in (1, 0):      # lone k <= 2 gates: direct kernel (default) and tile kernel
                hb.lib.hq_set_tuning(-1, -1, use_direct)
                st = hb.DeviceState(n, ctype).upload(psi)
                out = st.apply(U, pos).download()
                assert np.abs(out - ref).max() <= TOL[ctype], (k, pos, use_direct)
            hb.lib.hq_set_tuning(-1, -1, 1)
            if k <= 3:
                out = hb.DeviceState(n, ctype).upload(psi).apply(U, pos, direct=True).download()
                assert np.abs(out - ref).max() <= TOL[ctype], ("direct", k, pos)


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_tensor_core_gates_vs_oracle(hb, oracle, c_oracle, ctype):
    """mma.sync path (3xTF32 / FP64), k = 2..6, targets on the lowest bits (incl. amplitude bit 0, i.e.
    inside the 16-byte unit for complex64), the highest bits and random bits; two unmerged gates per pass
    so that k = 2 takes the tensor-core path too; full-size and small tiles."""
    rng = np.random.default_rng(77)
    for n in (14, 9):
        psi = _rand_state(rng, n, ctype)
        for k in range(2, 7):
            for variant in range(4):
                if variant == 0:
                    pos = list(range(k))
                elif variant == 1:
                    pos = list(range(n - k, n))
                else:
                    pos = [int(x) for x in rng.permutation(n)[:k]]
                pos = [int(x) for x in rng.permutation(pos)]
                gates = [(_haar(rng, k), pos), (_haar(rng, k), pos[::-1])]
                ref = oracle.evolve_oracle(psi, [(U.astype(ctype), p) for U, p in gates], c_oracle)
                for T in (0, 10):
                    plan = hb.Plan(gates, n, ctype, hb.PlanOptions(T, 1, 1, 0, 0, 0, -1, 1, 2))
                    assert plan.n_passes == 1 and plan.n_kernel_gates == 2
                    st = hb.DeviceState(n, ctype).upload(psi)
                    plan.run(st)
                    assert np.abs(st.download() - ref).max() <= TOL[ctype], (n, k, pos, T)


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_plan_variants_vs_oracle(hb, oracle, c_oracle, ctype):
    from hybridq_b200.circuits import matching_circuit, to_positions
    rng = np.random.default_rng(22)
    n = 18
    gates = matching_circuit(n, depth=8, seed=18)
    lowered, _ = to_positions(gates)
    # add a few wide gates (k = 3..6) so every kernel class runs inside fused passes
    for k in (3, 4, 5, 6):
        lowered.insert(int(rng.integers(0, len(lowered))), (_haar(rng, k), [int(x) for x in rng.permutation(n)[:k]]))
    psi = _rand_state(rng, n, ctype)
    ref = oracle.evolve_oracle(psi, [(U.astype(ctype), p) for U, p in lowered], c_oracle)
    try:
        for nbuf in (1, 2):
            hb.lib.hq_set_tuning(nbuf, 0, -1)
            for opts in (None, hb.PlanOptions(10, 3, 1, 0, 0), hb.PlanOptions(13, 5, 1, 0, 0),
                         hb.PlanOptions(12, 5, 0, 0, 0), hb.PlanOptions(11, 1, 1, 3, 0),
                         hb.PlanOptions(12, 5, 1, 0, 0, 0, -1), hb.PlanOptions(12, 4, 1, 0, 0, 2, 0),
                         hb.PlanOptions(13, 5, 1, 0, 0, 3, 24), hb.PlanOptions(mma_min_k=2), hb.PlanOptions(mma_min_k=0),
                         hb.PlanOptions(13, 5, 1, 0, 0, 4, -1, 1, 2), hb.PlanOptions(11, 2, 1, 0, 0, 3, -1, 1, 2),
                         hb.PlanOptions(12, 5, 1, 0, 0, 0, -1, 1, 2)):
                plan = hb.Plan(lowered, n, ctype, opts)
                st = hb.DeviceState(n, ctype).upload(psi)
                plan.run(st)
                out = st.download()
                assert plan.n_gates == len(lowered)
                assert np.abs(out - ref).max() <= TOL[ctype], (nbuf, opts and opts.tile_bits)
    finally:
        hb.lib.hq_set_tuning(0, 0, 1)


def test_simulate_golden(hb, golden):
    from hybridq_b200.circuits import GateApply
    z = golden["simulate"]
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"s{i}_ctype"])
        gq, n = golden_gates(z, f"s{i}")
        gates = [GateApply(U, tuple(int(x) for x in q)) for U, q in gq]     # qubit label = sorted index
        init = initial_from(z, f"s{i}_init", n, ctype)
        init_arg = init if isinstance(init, str) else init.reshape((2,) * n)
        out, info = hb.simulate(gates, initial_state=init_arg, complex_type=ctype, return_info=True)
        assert out.shape == (2,) * n and out.dtype == np.dtype(ctype)
        assert np.abs(out.reshape(-1) - z[f"s{i}_out"]).max() <= 4 * TOL[ctype], (i, str(z[f"s{i}_tag"]))
        assert info["n_passes"] < info["n_gate_applies"]


def test_functional_gates_golden(hb, golden):
    """Projection and Measure run on the device inside simulate() (no D2H/H2D round trip) and reproduce the
    reference's results, draw included (numpy's generator seeded as in the golden run)."""
    from hybridq_b200.circuits import GateApply, ProjectionApply, MeasureApply
    z = golden["functional"]
    n = int(z["n_qubits"])
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"f{i}_ctype"])
        circ = [GateApply(p, tuple(q)) if kind == "U" else (ProjectionApply(tuple(q), p) if kind == "P" else MeasureApply(tuple(q)))
                for kind, p, q in functional_items(z, i)]
        np.random.seed(int(z[f"f{i}_seed"]))
        launches0 = hb.lib.hq_launch_count()
        out = hb.simulate(circ, initial_state="+" * n, complex_type=ctype)
        assert np.abs(out.reshape(-1) - z[f"f{i}_out"]).max() <= 4 * TOL[ctype], i
        assert hb.lib.hq_launch_count() > launches0
    # the marginal of a product state
    st = hb.DeviceState(n, "complex128").init_product("+" * (n - 2) + "01")
    m = st.marginal([0, 1, 5]).sum(axis=1)                  # bit 0 = 1, bit 1 = 0, bit 5 = +
    want = np.zeros(8)
    want[0b001] = want[0b101] = 0.5
    assert np.abs(m - want).max() < 1e-12


def test_expectation_value_golden(hb, golden):
    """hybridq_b200.expectation_value against the reference's own expectation_value results."""
    from hybridq_b200.circuits import GateApply
    z = golden["expectation"]
    n = int(z["n_qubits"])
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"e{i}_ctype"])
        op = [GateApply(z[f"e{i}_g{j}_U"], tuple(int(x) for x in z[f"e{i}_g{j}_q"])) for j in range(int(z[f"e{i}_ngates"]))]
        state = z[f"e{i}_state"].astype(ctype).reshape((2,) * n)
        val = hb.expectation_value(state=state, op=op, qubits_order=list(range(n)), complex_type=ctype)
        assert abs(complex(val) - complex(z[f"e{i}_value"])) <= (2e-6 if ctype == "complex64" else 1e-12), i
    with pytest.raises(ValueError):
        hb.expectation_value(state=state, op=op, qubits_order=list(range(n - 1)), complex_type=ctype)
    with pytest.raises(ValueError):
        hb.expectation_value(state=state, op=[GateApply(np.eye(2), (99,))], qubits_order=list(range(n)))


def test_dm_golden(hb, golden):
    """Config 5 path: the lowered super-operator circuit (non-unitary 4x4 / 16x16 matrices)."""
    from hybridq_b200.circuits import GateApply
    z = golden["dm"]
    n = int(z["n_super"])
    for i in range(int(z["n_cases"])):
        ctype = str(z[f"m{i}_ctype"])
        gates = [GateApply(z[f"m{i}_g{j}_U"], tuple(int(x) for x in z[f"m{i}_g{j}_q"]))
                 for j in range(int(z[f"m{i}_ngates"]))]
        init = z[f"m{i}_init"]
        init_arg = str(init) if init.dtype.kind in "US" else init.astype(ctype).reshape((2,) * n)
        out = hb.simulate(gates, initial_state=init_arg, complex_type=ctype)
        assert np.abs(out.reshape(-1) - z[f"m{i}_out"]).max() <= TOL[ctype]
        rho = out.reshape(2 ** (n // 2), 2 ** (n // 2))
        assert abs(np.trace(rho) - 1) < 1e-5 and np.abs(rho - rho.conj().T).max() < 1e-5


def test_dot_and_transpose_golden(hb, golden):
    z = golden["dot"]
    n = int(z["n"])
    for i in range(int(z["n_cases"])):
        psi, U, axes, ref = z[f"d{i}_psi"], z[f"d{i}_U"], z[f"d{i}_axes"], z[f"d{i}_out"]
        ctype = str(psi.dtype)
        out = hb.dot(U, psi.reshape((2,) * n), axes_b=axes)
        assert np.abs(out.reshape(-1) - ref).max() <= TOL[ctype], i
        planes = np.array([psi.real.reshape((2,) * n), psi.imag.reshape((2,) * n)])
        out2 = hb.dot(U, planes, axes_b=axes, b_as_complex_array=True)
        assert np.abs((out2[0] + 1j * out2[1]).reshape(-1) - ref).max() <= TOL[ctype], i
    z = golden["transpose"]
    n = int(z["n"])
    for i in range(int(z["n_cases"])):
        dt = str(z[f"t{i}_dtype"])
        a = (np.arange(2 ** n, dtype=np.int64) * 5 + 1).astype(dt).reshape((2,) * n)
        out = hb.transpose(a, z[f"t{i}_axes"])
        assert np.array_equal(out.reshape(-1), z[f"t{i}_out"]), i


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_swap_dev_and_permute_bits(hb, oracle, ctype):
    rng = np.random.default_rng(23)
    n = 16
    psi = _rand_state(rng, n, ctype)
    for trial in range(6):
        m = int(rng.integers(2, 11))
        pos = rng.permutation(m)
        out = hb.DeviceState(n, ctype).upload(psi).swap(pos).download()
        assert np.array_equal(out, oracle.numpy_swap(psi, pos)), (trial, "swap")      # bit-exact
        perm = rng.permutation(n)
        out = hb.DeviceState(n, ctype).upload(psi).permute_bits(perm).download()
        assert np.array_equal(out, oracle.numpy_swap(psi, perm)), (trial, "perm")


def test_init_product_and_reductions(hb):
    n = 14
    for ctype in ("complex64", "complex128"):
        for spec in ("0" * n, "+-01" * 3 + "1+", "1" * n):
            st = hb.DeviceState(n, ctype).init_product(spec)
            ref = product_state(spec, ctype)
            assert np.abs(st.download() - ref).max() < (1e-6 if ctype == "complex64" else 1e-14)
            assert abs(st.norm2() - 1) < 1e-5
        a = hb.DeviceState(n, ctype).init_random(seed=5)
        b = hb.DeviceState(n, ctype).init_random(seed=6)
        ha, hbb = a.download(), b.download()
        assert abs(a.norm2() - 1) < 1e-6
        assert abs(a.vdot(b) - np.vdot(ha.astype(np.complex128), hbb.astype(np.complex128))) < 1e-6


def test_functional_gate_round_trip(hb, oracle, c_oracle):
    """FunctionalGate contract (simulation.py:525-554): gets split planes + order on the host."""
    from hybridq_b200.circuits import matching_circuit, to_positions

    class Project0:                      # zero the amplitudes where the first qubit is 1
        qubits = (0,)

        def apply(self, psi, order):
            assert psi.shape[0] == 2 and order[0] == 0
            psi = psi.copy()
            psi[:, 1] = 0
            return psi, order

    n = 12
    gates = matching_circuit(n, depth=3, seed=3)
    mixed = gates[:10] + [Project0()] + gates[10:]
    out = hb.simulate(mixed, initial_state="+" * n, complex_type="complex128")
    lowered, _ = to_positions(gates)
    psi = product_state("+" * n, "complex128")
    psi = oracle.evolve_oracle(psi, lowered[:10], c_oracle)
    psi = psi.reshape((2,) * n).copy()
    psi[1] = 0
    psi = oracle.evolve_oracle(psi.reshape(-1), lowered[10:], c_oracle)
    assert np.abs(out.reshape(-1) - psi).max() <= 1e-12


# ------------------------------------------------------------------ BASELINE sizes: properties
@pytest.mark.parametrize("n,ctype", [(30, "complex64"), (29, "complex128")])
def test_full_size_round_trip_and_norm(hb, n, ctype):
    """No CPU oracle finishes in seconds at this size: check size-independent properties.
    U then U^dagger (reversed circuit) must return the initial state; the norm is preserved;
    a fused plan and a one-gate-per-pass plan must agree."""
    from hybridq_b200.circuits import matching_circuit, ksweep_circuit, to_positions
    gates = matching_circuit(n, depth=2, seed=n) + ksweep_circuit(n, 3, 2) + ksweep_circuit(n, 5, 1)
    lowered, _ = to_positions(gates)
    inverse = [(U.conj().T, p) for U, p in reversed(lowered)]
    st = hb.DeviceState(n, ctype).init_random(seed=n)
    ref = st.copy()
    fwd = hb.Plan(lowered, n, ctype)
    bwd = hb.Plan(inverse, n, ctype, hb.PlanOptions(0, -1, 0, 0, 0))      # unfused on the way back
    fwd.run(st)
    n2 = st.norm2()
    assert abs(n2 - 1) < (1e-4 if ctype == "complex64" else 1e-10)
    ov_mid = abs(ref.vdot(st))
    assert ov_mid < 0.5                       # the circuit really changed the state
    bwd.run(st)
    ov = ref.vdot(st)
    assert abs(ov - 1) < (1e-4 if ctype == "complex64" else 1e-10)
    # element-wise: max-abs deviation from the initial state
    import torch
    err = float((st.tensor - ref.tensor).abs().max())
    assert err <= TOL[ctype], err


# ------------------------------------------------------------------ wide Measure / Projection
def test_measure_and_projection_on_more_than_ten_qubits(hb, oracle):
    """The reference's Measure / Projection take any number of qubits (gate/measure.py:77, gate/projection.py:72);
    measuring every qubit is the usual sampling idiom.  k = 14 > 10 goes through the global-memory histogram and
    reproduces the reference's draw (same numpy generator state); k = 12 projection through the conditional sum."""
    from hybridq_b200.circuits import matching_circuit, to_positions, MeasureApply, ProjectionApply, random_state
    n = 14
    gates = matching_circuit(n, depth=3, seed=14)
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    for ctype in ("complex64", "complex128"):
        psi0 = random_state(n, ctype, seed=2)
        mid = oracle.evolve_oracle(psi0, [(U.astype(ctype), p) for U, p in lowered])
        # Measure over all qubits, given in a scrambled order
        order = [int(x) for x in np.random.default_rng(5).permutation(n)]
        np.random.seed(1234)
        want, s = oracle.numpy_measure(mid, [n - 1 - q for q in order])
        np.random.seed(1234)
        out = hb.simulate(gates + [MeasureApply(tuple(order))], initial_state=psi0.reshape((2,) * n), complex_type=ctype)
        assert np.abs(out.reshape(-1) - want).max() <= TOL[ctype]
        assert np.count_nonzero(out) == 1
        # Projection of 12 of the 14 qubits
        qs = order[:12]
        bits = "".join(str((i * 7 + 3) % 2) for i in range(12))
        want = oracle.numpy_project(mid, [n - 1 - q for q in qs], [int(b) for b in bits])
        out = hb.simulate(gates + [ProjectionApply(tuple(qs), bits)], initial_state=psi0.reshape((2,) * n), complex_type=ctype)
        assert np.abs(out.reshape(-1) - want).max() <= 4 * TOL[ctype]


def test_measure_more_than_24_qubits_samples_in_chunks(hb):
    """26 measured qubits: the outcome is drawn group by group from conditional marginals; the result must be
    the basis state of the drawn outcome, and over a product state with known single-qubit probabilities the
    drawn bits must be the certain ones."""
    from hybridq_b200.circuits import MeasureApply, GateApply
    n = 26
    spec = "01" * 13                                   # qubit q (label q = string position) is |0> or |1>
    H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    np.random.seed(7)
    circ = [GateApply(H, (3,)), GateApply(H, (20,)), MeasureApply(tuple(range(n)))]
    st = hb.simulate(circ, initial_state=spec, complex_type="complex64", return_numpy_array=False)
    assert abs(st.norm2() - 1) < 1e-6
    out = st.download()
    idx = np.flatnonzero(out)
    assert idx.size == 1 and abs(abs(out[idx[0]]) - 1) < 1e-6
    bits = format(int(idx[0]), f"0{n}b")               # string position q <-> qubit label q (MSB first)
    for q in range(n):
        if q not in (3, 20):
            assert bits[q] == spec[q], q


def test_state_on_a_device_that_is_not_current(hb, oracle, c_oracle):
    """DeviceState(device=1) while cuda:0 is current must launch on cuda:1."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from hybridq_b200.circuits import matching_circuit, to_positions, random_state
    n = 16
    lowered, _ = to_positions(matching_circuit(n, depth=4, seed=1), qubits=list(range(n)))
    psi0 = random_state(n, "complex64", seed=3)
    ref = oracle.evolve_oracle(psi0, [(U.astype("complex64"), p) for U, p in lowered], c_oracle)
    torch.cuda.set_device(0)
    st = hb.DeviceState(n, "complex64", device=1).upload(psi0)
    hb.Plan(lowered, n, "complex64").run(st)
    assert st.tensor.device.index == 1 and torch.cuda.current_device() == 0
    assert np.abs(st.download() - ref).max() <= TOL["complex64"]
    out = hb.simulate(matching_circuit(n, depth=4, seed=1), initial_state=psi0.reshape((2,) * n), device=1)
    assert np.abs(out.reshape(-1) - ref).max() <= TOL["complex64"]


# ------------------------------------------------------------------ checkpoint and sampling
def test_checkpoint_round_trip_and_sampling(hb, tmp_path):
    """dump / load through a small pinned staging buffer (several chunks) is bit-exact; a checkpoint of another
    size or precision is refused; sample() draws from |psi|^2 without touching the state."""
    from hybridq_b200.circuits import matching_circuit
    n = 18
    st = hb.simulate(matching_circuit(n, depth=4, seed=18), initial_state="0" * n, complex_type="complex64",
                     return_numpy_array=False)
    ref = st.download()
    st.dump(tmp_path / "ckpt.bin", chunk_bytes=1 << 18)
    back = hb.DeviceState(n, "complex64").load(tmp_path / "ckpt.bin", chunk_bytes=3 << 17)
    assert np.array_equal(back.download(), ref)
    with pytest.raises(ValueError):
        hb.DeviceState(n, "complex128").load(tmp_path / "ckpt.bin")
    with pytest.raises(ValueError):
        hb.DeviceState(n - 1, "complex64").load(tmp_path / "ckpt.bin")
    # sampling: a state with three populated basis states of known weights
    psi = np.zeros(2 ** n, dtype=np.complex64)
    psi[5], psi[70000], psi[2 ** n - 1] = np.sqrt(0.5), 1j * np.sqrt(0.3), -np.sqrt(0.2)
    st = hb.DeviceState(n, "complex64").upload(psi)
    draws = st.sample(20000, seed=1, block_bits=10)
    assert set(np.unique(draws)) == {5, 70000, 2 ** n - 1}
    freq = np.array([(draws == i).mean() for i in (5, 70000, 2 ** n - 1)])
    assert np.abs(freq - [0.5, 0.3, 0.2]).max() < 0.02
    assert np.array_equal(st.download(), psi)              # not collapsed
    # and of an evolved state: empirical distribution of the top 4 bits vs the exact marginal
    draws = back.sample(50000, seed=2)
    top = np.bincount(draws >> (n - 4), minlength=16) / 50000
    want = back.marginal(list(range(n - 4, n))).sum(axis=1)
    assert np.abs(top - want).max() < 0.02


def test_simulate_with_pinned_host_arrays_folds_the_transfers(hb, oracle, c_oracle):
    """hq_plan_run_io: with pinned `initial_state` / `out` the first pass reads the host array and the last pass
    writes the host array; same result as the copy-in / copy-out path and as the oracle, for one pass and many."""
    import torch
    from hybridq_b200.circuits import matching_circuit, to_positions, random_state
    for n, depth, ctype in ((17, 6, "complex64"), (12, 2, "complex128"), (18, 3, "complex64")):
        gates = matching_circuit(n, depth=depth, seed=n)     # n = 12 complex128 is one tile: a single pass carries
                                                             # the upload and the download
        lowered, _ = to_positions(gates, qubits=list(range(n)))
        psi0 = random_state(n, ctype, seed=4)
        ref = oracle.evolve_oracle(psi0, [(U.astype(ctype), p) for U, p in lowered], c_oracle)
        tdt = torch.complex64 if ctype == "complex64" else torch.complex128
        h_in = torch.empty(2 ** n, dtype=tdt, pin_memory=True)
        h_out = torch.empty(2 ** n, dtype=tdt, pin_memory=True)
        h_in.numpy()[:] = psi0
        h_out.zero_()
        psi, info = hb.simulate(gates, initial_state=h_in.numpy().reshape((2,) * n), complex_type=ctype,
                                out=h_out.numpy(), return_info=True)
        assert info["transfers folded into passes"] == {"upload": True, "download": True}
        assert np.abs(h_out.numpy() - ref).max() <= TOL[ctype]
        assert np.shares_memory(psi, h_out.numpy())
        assert np.array_equal(h_in.numpy(), psi0)            # the input is left alone
        # pageable arrays take the ordinary path and give the same numbers
        plain, info2 = hb.simulate(gates, initial_state=psi0.reshape((2,) * n), complex_type=ctype, return_info=True)
        assert info2["transfers folded into passes"] == {"upload": False, "download": False}
        assert np.abs(plain.reshape(-1) - h_out.numpy()).max() <= TOL[ctype]


def test_lone_scalar_plus_rank_one_gate(hb, oracle, c_oracle):
    """A pass holding ONE scalar + rank-one k = 4 gate (fuse = 0, or simply a one-gate circuit) must take the tile
    kernel, not the k <= 3 direct kernel (regression: the header's kernel class is 3 for such a gate)."""
    rng = np.random.default_rng(12)
    n = 14
    u = rng.standard_normal(16) + 1j * rng.standard_normal(16)
    U = 0.95 * np.eye(16) + 0.01 * np.outer(u, u.conj())
    for ctype in ("complex64", "complex128"):
        psi = _rand_state(rng, n, ctype)
        ref = oracle.evolve_oracle(psi, [(U.astype(ctype), [2, 5, 8, 11])], c_oracle)
        for opts in (None, hb.PlanOptions(fuse=0)):
            st = hb.DeviceState(n, ctype).upload(psi)
            plan = hb.Plan([(U, [2, 5, 8, 11])], n, ctype, opts)
            assert plan.arithmetic() == {"k4": {"scalar_plus_rank_one": 1}}
            plan.run(st)
            assert np.abs(st.download() - ref).max() <= TOL[ctype]


# ------------------------------------------------------------------ tcgen05 lone-gate kernel (hq_umma.cuh)
@pytest.mark.parametrize("n,pos", [
    (12, [1, 4, 7, 9, 11]), (12, [0, 1, 2, 3, 4]), (13, [0, 3, 8, 10, 12]), (16, [2, 5, 6, 11, 15]),
    (16, [1, 2, 3, 14, 15]), (20, [0, 1, 12, 17, 19]), (22, [3, 7, 12, 20, 21]),
    (11, [0, 1, 2, 3]), (11, [2, 6, 9, 10]), (14, [0, 5, 6, 13]), (18, [1, 2, 16, 17]), (21, [4, 9, 13, 20]),
    (13, [0, 1, 2, 3, 4, 5]), (13, [1, 3, 5, 7, 9, 12]), (15, [0, 2, 7, 8, 13, 14]), (19, [3, 4, 10, 11, 17, 18]),
    (22, [0, 1, 8, 15, 20, 21]),
])
def test_tcgen05_lone_gate_parity(hb, oracle, c_oracle, n, pos):
    """One dense complex64 k = 4 / 5 / 6 gate = one pass on the tcgen05 kernel; checked against the oracle, against the
    mma.sync tile-kernel path it replaces, and (launch counter) that it really ran."""
    rng = np.random.default_rng(100 * n + len(pos) + pos[0])
    k = len(pos)
    U = _haar(rng, k)
    psi = _rand_state(rng, n, "complex64")
    perm = rng.permutation(k)                           # matrix bit order need not be ascending
    gate = (U, [pos[i] for i in perm])
    ref = oracle.evolve_oracle(psi, [(U.astype("complex64"), gate[1])], c_oracle)
    plan = hb.Plan([gate], n, "complex64")
    assert plan.n_umma_passes == 1
    before = hb.lib.hq_umma_launch_count()
    st = hb.DeviceState(n, "complex64").upload(psi)
    plan.run(st)
    out = st.download()
    assert hb.lib.hq_umma_launch_count() == before + 1
    assert np.abs(out - ref).max() <= TOL["complex64"]
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()         # near fp32 accuracy, not merely 1e-6 absolute
    old = hb.lib.hq_set_umma(0)
    try:
        st2 = hb.DeviceState(n, "complex64").upload(psi)
        plan.run(st2)
        assert hb.lib.hq_umma_launch_count() == before + 1
        assert np.abs(st2.download() - out).max() <= 2e-6 * np.abs(ref).max()
    finally:
        hb.lib.hq_set_umma(old)


def test_tcgen05_path_scope(hb):
    """Only complex64 passes of one dense k = 4 .. 6 matrix on >= k + 7 qubits qualify; everything else keeps its path."""
    rng = np.random.default_rng(5)
    U5, U4, U3 = _haar(rng, 5), _haar(rng, 4), _haar(rng, 3)
    U6 = _haar(rng, 6)
    assert hb.Plan([(U6, [0, 1, 2, 3, 4, 5])], 13, "complex64").n_umma_passes == 1
    assert hb.Plan([(U6, [0, 1, 2, 3, 4, 5])], 12, "complex64").n_umma_passes == 0
    assert hb.Plan([(U5, [0, 1, 2, 3, 4])], 12, "complex128").n_umma_passes == 0
    assert hb.Plan([(U5, [0, 1, 2, 3, 4])], 11, "complex64").n_umma_passes == 0        # fewer than 128 groups
    assert hb.Plan([(U3, [0, 1, 2])], 12, "complex64").n_umma_passes == 0
    assert hb.Plan([(U4, [0, 1, 2, 3]), (U5, [4, 5, 6, 7, 8])], 14, "complex64", hb.PlanOptions(fuse=0)).n_umma_passes == 2
    assert hb.Plan([(U4, [0, 1, 2, 3])], 14, "complex64", hb.PlanOptions(mma_min_k=0)).n_umma_passes == 0


def test_tcgen05_norm_and_error_over_many_gates(hb):
    """300 random 5-qubit unitaries one after another (each its own tcgen05 pass).  Tensor cores add into their fp32
    accumulator with truncation, so some norm loss per gate is inherent to both tensor-core paths (measured on B200,
    profiles/r02/umma_drift.jsonl: -2.0e-5 here, -1.1e-5 on the mma.sync path, after 300 gates); the kernel keeps it
    there by giving the big hi * hi products short accumulator chains (one chain for everything: -2.0e-4).  The state
    must stay within the complex64 tolerance of the complex128 run."""
    rng = np.random.default_rng(77)
    n = 16
    gates = []
    for _ in range(300):
        pos = sorted(rng.choice(n, size=5, replace=False).tolist())
        gates.append((_haar(rng, 5), pos))
    psi = _rand_state(rng, n, "complex128")
    opts = hb.PlanOptions(fuse=0)
    plan32 = hb.Plan(gates, n, "complex64", opts)
    assert plan32.n_umma_passes == 300
    st32 = hb.DeviceState(n, "complex64").upload(psi.astype("complex64"))
    before = hb.lib.hq_umma_launch_count()
    plan32.run(st32)
    assert hb.lib.hq_umma_launch_count() == before + 300
    st64 = hb.DeviceState(n, "complex128").upload(psi)
    hb.Plan(gates, n, "complex128", opts).run(st64)
    a, b = st32.download(), st64.download()
    drift = np.linalg.norm(a.astype("complex128")) - 1.0
    assert abs(drift) < 5e-5, drift
    assert np.abs(a - b).max() <= TOL["complex64"]
    old = hb.lib.hq_set_umma(0)
    try:
        st = hb.DeviceState(n, "complex64").upload(psi.astype("complex64"))
        plan32.run(st)
        drift_mma_sync = np.linalg.norm(st.download().astype("complex128")) - 1.0
    finally:
        hb.lib.hq_set_umma(old)
    assert abs(drift) < 3.0 * abs(drift_mma_sync) + 1e-6, (drift, drift_mma_sync)


def test_big_gates_inside_a_circuit_run_as_solo_tcgen05_passes(hb, oracle, c_oracle):
    """A complex64 circuit of 1-/2-qubit gates with dense 4-, 5- and 6-qubit gates in between: the planner gives each
    big gate a pass of its own (tcgen05 kernel), multiplies the gates acting inside its qubits into it, and fuses the
    rest around it; result against the oracle, and against the complex128 plan (all tile passes)."""
    rng = np.random.default_rng(2024)
    n = 16
    gates = []
    big = 0
    for layer in range(12):
        for _ in range(6):
            k = int(rng.integers(1, 3))
            gates.append((_haar(rng, k), [int(x) for x in rng.permutation(n)[:k]]))
        if layer % 2 == 0:
            k = 4 + (layer // 2) % 3
            pos = [int(x) for x in rng.permutation(n)[:k]]
            gates.append((_haar(rng, k), pos))
            gates.append((_haar(rng, 2), pos[1:3]))              # acts inside the big gate: absorbed
            big += 1
    psi = _rand_state(rng, n, "complex64")
    ref = oracle.evolve_oracle(psi, [(U.astype("complex64"), p) for U, p in gates], c_oracle)
    plan = hb.Plan(gates, n, "complex64")
    assert plan.n_umma_passes == big == 6
    assert plan.n_passes < len(gates) // 3                        # the small gates are still fused
    before = hb.lib.hq_umma_launch_count()
    st = hb.DeviceState(n, "complex64").upload(psi)
    plan.run(st)
    out = st.download()
    assert hb.lib.hq_umma_launch_count() == before + big
    assert np.abs(out - ref).max() <= TOL["complex64"]
    st128 = hb.DeviceState(n, "complex128").upload(psi.astype("complex128"))
    p128 = hb.Plan(gates, n, "complex128")
    assert p128.n_umma_passes == 0
    p128.run(st128)
    assert np.abs(st128.download() - out).max() <= TOL["complex64"]


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_lone_small_gates_take_the_direct_kernel(hb, oracle, c_oracle, ctype):
    """A pass made of one gate with k <= 3 runs on the shared-memory-free direct kernel at copy bandwidth, in both
    precisions (regression: complex128 k = 2, 3 gates carry the row-pair kind and were sent to the tile kernel --
    config 3's lone k = 3 row dropped from 20.5 to 7.9 gate-applies/s before this was caught)."""
    rng = np.random.default_rng(41)
    n = 14
    for k, pos in ((1, [6]), (2, [0, 9]), (2, [3, 12]), (3, [1, 7, 13]), (3, [0, 4, 8])):
        U = _haar(rng, k)
        psi = _rand_state(rng, n, ctype)
        ref = oracle.evolve_oracle(psi, [(U.astype(ctype), pos)], c_oracle)
        before = hb.lib.hq_direct_launch_count()
        st = hb.DeviceState(n, ctype).upload(psi)
        hb.Plan([(U, pos)], n, ctype).run(st)
        assert hb.lib.hq_direct_launch_count() == before + 1, (ctype, k, pos)
        assert np.abs(st.download() - ref).max() <= TOL[ctype]
