"""GPU parity at BASELINE scale (VERDICT r01 "next round" item 1).

Every test compares the FULL final vector produced by the CUDA path (through the C ABI, default plan = the
fused / merged plan bench.py times) with the reference's own compiled C++ core (oracle/_ref/{wheel,avx2},
built from /root/reference by oracle/Makefile and shipped to the GPU box) driven by
``oracle.evolve_ref`` exactly like the reference's host loop (simulation.py:522-663).
Tolerances: max-abs 1e-6 (complex64), 1e-12 (complex128) on normalised states (BASELINE.json north_star).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import TOL, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hybridq_b200
    return hybridq_b200


@pytest.fixture(scope="module")
def ref_core(oracle):
    """The reference's compiled core: upstream's wheel binary when present (fastest), else the local AVX2 build."""
    for variant in ("wheel", "avx2"):
        if oracle.RefCore.available(variant):
            return oracle.RefCore(variant)
    pytest.skip("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")


def _device_max_abs_diff(state, ref_host):
    """max |state - ref| with the comparison done on the device in chunks (test infrastructure: torch is only
    the checker here, the state was produced by the C-ABI kernels)."""
    import torch
    ref_host = np.ascontiguousarray(ref_host).reshape(-1)
    t = state.tensor
    worst = 0.0
    chunk = 1 << 26
    for a in range(0, t.numel(), chunk):
        r = torch.from_numpy(ref_host[a:a + chunk]).to(t.device)
        worst = max(worst, float((t[a:a + chunk] - r).abs().max()))
    return worst


# ------------------------------------------------------------------------------ config 2: the bench circuit
@pytest.mark.parametrize("n,ctype", [(24, "complex64"), (24, "complex128"), (26, "complex64"), (26, "complex128")])
def test_bench_circuit_full_vector_vs_reference_core(hb, oracle, ref_core, n, ctype):
    """matching_circuit(n, 20, seed=n) -- the generator bench.py uses -- with the default plan, n = 24 / 26."""
    from hybridq_b200.circuits import matching_circuit, to_positions
    lowered, _ = to_positions(matching_circuit(n, depth=20, seed=n), qubits=list(range(n)))
    st = hb.DeviceState(n, ctype).init_random(seed=n)
    psi0 = st.download()
    plan = hb.Plan(lowered, n, ctype)
    assert plan.n_passes < plan.n_gates            # fused
    plan.run(st)
    ref = oracle.evolve_ref(psi0, [(U.astype(ctype), p) for U, p in lowered], ref_core)
    err = _device_max_abs_diff(st, ref)
    assert err <= TOL[ctype], (n, ctype, err)
    assert abs(st.norm2() - 1) < (2e-6 if ctype == "complex64" else 1e-12)


def test_bench_circuit_n30_complex64_vs_reference_core(hb, oracle, ref_core):
    """BASELINE config[1] itself: n = 30, depth 20, complex64, all 471 gate-applies, full 2^30 vector against
    the reference core on the host (about a minute of CPU time; ~24 GB of host memory)."""
    from hybridq_b200.circuits import matching_circuit, to_positions
    n, ctype = 30, "complex64"
    try:
        import psutil
        if psutil.virtual_memory().available < 36 * 2 ** 30:
            pytest.skip("needs ~32 GiB of free host memory")
    except ImportError:
        pass
    lowered, _ = to_positions(matching_circuit(n, depth=20, seed=n), qubits=list(range(n)))
    st = hb.DeviceState(n, ctype).init_random(seed=n)
    psi0 = st.download()
    plan = hb.Plan(lowered, n, ctype)
    plan.run(st)
    st.sync()
    ref = oracle.evolve_ref(psi0, [(U.astype(ctype), p) for U, p in lowered], ref_core)
    del psi0
    err = _device_max_abs_diff(st, ref)
    assert err <= TOL[ctype], err
    assert abs(st.norm2() - 1) < 2e-6


# ------------------------------------------------------------------------------ config 3: k sweep, complex128
def test_config3_k_sweep_n28_complex128_vs_reference_core(hb, oracle, ref_core):
    """Config 3's k = 1..6 sweep (ksweep_circuit, seeds 331..336) at n = 28 complex128 (4 GiB state): every k as
    lone gates (one launch per gate: direct kernel / tile kernel / tensor-core path) and as one fused plan,
    both against the reference core.  Gate counts are trimmed for k >= 4 to bound the CPU time."""
    from hybridq_b200.circuits import ksweep_circuit, to_positions
    n, ctype = 28, "complex128"
    counts = {1: 6, 2: 6, 3: 4, 4: 3, 5: 2, 6: 2}
    lowered = []
    for k, cnt in counts.items():
        lowered += to_positions(ksweep_circuit(n, k, n_gates=cnt), qubits=list(range(n)))[0]
    st = hb.DeviceState(n, ctype).init_random(seed=3)
    psi0 = st.download()
    lone = st.copy()
    hb.Plan(lowered, n, ctype).run(st)                                         # fused
    hb.Plan(lowered, n, ctype, hb.PlanOptions(fuse=0)).run(lone)               # one pass per gate
    ref = oracle.evolve_ref(psi0, [(U.astype(ctype), p) for U, p in lowered], ref_core)
    del psi0
    assert _device_max_abs_diff(st, ref) <= TOL[ctype]
    assert _device_max_abs_diff(lone, ref) <= TOL[ctype]


def test_big_gates_n28_complex64_tcgen05_vs_reference_core(hb, oracle, ref_core):
    """SURVEY 8(a2) at scale on the tcgen05 path: dense k = 4, 5, 6 gates (ksweep_circuit, the config-3 generator) on a
    2^28 complex64 state, one pass per gate = one launch of hq_umma_gate_kernel each, plus gates on the lowest bits
    (memory-order lane map, 128-bit units); FULL vector against the reference's compiled core."""
    from hybridq_b200.circuits import ksweep_circuit, to_positions, haar_unitary
    n, ctype = 28, "complex64"
    lowered = []
    for k, cnt in ((4, 3), (5, 3), (6, 2)):
        lowered += to_positions(ksweep_circuit(n, k, n_gates=cnt), qubits=list(range(n)))[0]
    rng = np.random.default_rng(28)
    for pos in ([0, 1, 2, 3], [0, 1, 2, 3, 4], [1, 2, 3, 4, 5, 27], [0, 2, 3, 26, 27]):
        lowered.append((haar_unitary(2 ** len(pos), rng), pos))
    st = hb.DeviceState(n, ctype).init_random(seed=5)
    psi0 = st.download()
    plan = hb.Plan(lowered, n, ctype, hb.PlanOptions(fuse=0))
    assert plan.n_umma_passes == len(lowered) == 12
    before = hb.lib.hq_umma_launch_count()
    plan.run(st)
    st.sync()
    assert hb.lib.hq_umma_launch_count() == before + 12
    ref = oracle.evolve_ref(psi0, [(U.astype(ctype), p) for U, p in lowered], ref_core)
    del psi0
    err = _device_max_abs_diff(st, ref)
    assert err <= TOL[ctype], err
    assert err <= 1e-8, err                   # 2^-14 amplitudes: twelve gates stay at the 1e-9 level
    assert abs(st.norm2() - 1) < 2e-6


def test_big_gates_n30_complex64_round_trip(hb):
    """BASELINE size (n = 30, 8 GiB state), size-independent property: U then U^dagger on the tcgen05 path gives the
    state back (k = 4, 5, 6; spread targets and the lowest bits), and the norm is kept."""
    from hybridq_b200.circuits import haar_unitary
    import torch
    n, ctype = 30, "complex64"
    rng = np.random.default_rng(30)
    st = hb.DeviceState(n, ctype).init_random(seed=9)
    ref = st.copy()
    gates = []
    for pos in ([3, 9, 17, 29], [0, 1, 2, 3], [2, 8, 15, 22, 28], [0, 1, 2, 3, 4], [1, 6, 13, 19, 24, 29], [0, 1, 2, 3, 4, 5]):
        U = haar_unitary(2 ** len(pos), rng)
        gates += [(U, pos), (U.conj().T, pos)]
    plan = hb.Plan(gates, n, ctype, hb.PlanOptions(fuse=0, merge_max_k=0))
    assert plan.n_umma_passes == len(gates)
    before = hb.lib.hq_umma_launch_count()
    plan.run(st)
    st.sync()
    assert hb.lib.hq_umma_launch_count() == before + len(gates)
    worst = 0.0
    a, b = st.tensor, ref.tensor
    for lo in range(0, a.numel(), 1 << 27):
        worst = max(worst, float((a[lo:lo + (1 << 27)] - b[lo:lo + (1 << 27)]).abs().max()))
    assert worst <= 1e-8, worst                       # amplitudes ~ 3e-5: relative 3e-4 would already fail
    assert abs(st.norm2() - 1) < 2e-6


# ------------------------------------------------------------------------------ config 5: density matrices
def test_dm_10_and_12_qubits_vs_reference(hb, oracle, ref_core, golden):
    """2^20 and 2^24 superkets: the lowered circuits come from the reference's dm front-end
    (tests/golden/make_golden.py::make_dm_large); the full vector is compared with the reference core run here
    on the same lowered gates, and 8192 + 2^nq sampled amplitudes (incl. the whole diagonal), the trace and the
    squared norm with what the unmodified `hybridq.dm.circuit.simulation.simulate` returned."""
    z = golden["dm_large"]
    ctype = "complex64"
    for ci in range(int(z["n_cases"])):
        nq = int(z[f"L{ci}_nq"])
        n = 2 * nq
        gates = [(z[f"L{ci}_g{j}_U"], [n - 1 - int(x) for x in reversed(z[f"L{ci}_g{j}_q"])])
                 for j in range(int(z[f"L{ci}_ngates"]))]
        init = str(z[f"L{ci}_init"])
        st = hb.DeviceState(n, ctype).init_product(init)
        psi0 = st.download()
        plan = hb.Plan(gates, n, ctype)
        plan.run(st)
        ref = oracle.evolve_ref(psi0, [(U.astype(ctype), p) for U, p in gates], ref_core)
        assert _device_max_abs_diff(st, ref) <= TOL[ctype], nq
        out = st.download()
        assert np.abs(out[z[f"L{ci}_idx"]] - z[f"L{ci}_val"]).max() <= TOL[ctype], nq
        rho = out.reshape(2 ** nq, 2 ** nq)
        assert abs(np.trace(rho) - complex(z[f"L{ci}_trace"])) < 1e-5
        assert abs(np.vdot(out.astype(np.complex128), out.astype(np.complex128)).real - float(z[f"L{ci}_norm2"])) < 1e-5
        assert np.abs(rho - rho.conj().T).max() < 1e-6


# ------------------------------------------------------------------------------ depth (reference tests.py:2337)
def test_depth_600_n20_complex64_drift(hb, oracle, c_oracle, ref_core):
    """The reference's own large test runs depth-600 circuits at 16-22 qubits (tests/tests.py:2337
    test_simulation_4__simulation_large).  600 gate-applies at n = 20, complex64: max-abs vs the reference core
    in complex64 (<= 1e-6), vs the complex128 oracle (<= 2e-8: the reference's own c64-vs-c128 gap is a few 1e-9),
    and the norm drift after 600 gates (<= 2e-6; the reference's FMA arithmetic gives ~1e-7)."""
    from hybridq_b200.circuits import matching_circuit, to_positions, random_state
    n = 20
    lowered, _ = to_positions(matching_circuit(n, depth=42, seed=2337), qubits=list(range(n)))
    lowered = lowered[:600]
    assert len(lowered) == 600
    psi0 = random_state(n, "complex64", seed=7)
    st = hb.DeviceState(n, "complex64").upload(psi0)
    hb.Plan(lowered, n, "complex64").run(st)
    out = st.download()
    ref64 = oracle.evolve_ref(psi0, [(U.astype("complex64"), p) for U, p in lowered], ref_core)
    ref128 = oracle.evolve_oracle(psi0.astype("complex128"), [(U.astype("complex128"), p) for U, p in lowered], c_oracle)
    assert np.abs(out - ref64).max() <= 1e-6
    assert np.abs(out - ref128).max() <= 2e-8
    assert abs(st.norm2() - 1) <= 2e-6


# ------------------------------------------------------------------------------ multi-GPU: sharded vs one GPU
def test_sharded_two_gpus_vs_one_gpu_n26(hb):
    """torchrun with 2 ranks (one per GPU): the sharded evolution of the bench circuit at n = 26 must reproduce the
    single-GPU result amplitude for amplitude (tests/dist_gpu_worker.py does the comparison on rank 0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        str(Path(ROOT) / "tests" / "dist_gpu_worker.py"), "26"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_PARITY_OK" in r.stdout, r.stdout[-2000:]
