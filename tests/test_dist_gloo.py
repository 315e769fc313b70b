"""N > 1 path on CPU: world_size 2 and 4 over gloo.  Checks hybridq_b200.dist's schedule
(Belady remapping of rank bits, in-place rank-bit <-> local-bit swaps, restoration of the
canonical order) and the send/recv form of the exchange against a single-process oracle evolution.
(The fused NVLink form of the exchange is GPU-only: tests/test_gpu_ring.py plays it on one GPU,
tests/dist_gpu_worker.py on two.)"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _launch(world, n, ctype, seed, out, *extra):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world * 10 + seed),
           str(ROOT / "tests" / "dist_worker.py"), str(n), ctype, str(seed), str(out), *extra]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("world,seed", [(2, 1), (2, 2), (4, 3)])
def test_sharded_evolution_matches_oracle(tmp_path, oracle, world, seed):
    from hybridq_b200.circuits import sharded_circuit, matching_circuit, to_positions
    n, ctype = 10, "complex128"
    out = tmp_path / "res.npz"
    _launch(world, n, ctype, seed, out)
    z = np.load(out)
    g = int(np.log2(world))
    gates = sharded_circuit(n, g, depth=5, frac_global=0.3, seed=seed) if seed % 2 else matching_circuit(n, depth=4, seed=seed)
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    ref = oracle.evolve_oracle(z["psi"], lowered)
    assert np.abs(z["out"] - ref).max() < 1e-12
    assert abs(float(z["norm2"]) - 1) < 1e-12
    assert int(z["exchanges"]) >= 1 and int(z["exchanges"]) <= int(z["crossing"]) + 2


@pytest.mark.parametrize("world,seed", [(2, 5), (4, 7)])
def test_sharded_projection_and_measure(tmp_path, oracle, world, seed):
    """Projection and Measure on a sharded state (targets on rank bits and on local bits) between gate
    segments: same result as the single-process restatement with the same numpy seed."""
    from hybridq_b200.circuits import sharded_circuit, to_positions
    n, ctype = 10, "complex128"
    out = tmp_path / "res.npz"
    _launch(world, n, ctype, seed, out, "functional")
    z = np.load(out)
    g = int(np.log2(world))
    lowered, _ = to_positions(sharded_circuit(n, g, depth=5, frac_global=0.3, seed=seed), qubits=list(range(n)))
    cut = [len(lowered) // 3, 2 * len(lowered) // 3]
    psi = oracle.evolve_oracle(z["psi"], lowered[:cut[0]])
    psi = oracle.numpy_project(psi, [n - 1 - 0, n - 1 - (n - 2)], [1, 0])
    psi = oracle.evolve_oracle(psi, lowered[cut[0]:cut[1]])
    np.random.seed(seed)
    psi, _ = oracle.numpy_measure(psi, [n - 1 - q for q in (n - 1, 1, 4)])
    psi = oracle.evolve_oracle(psi, lowered[cut[1]:])
    assert np.abs(z["out"] - psi).max() < 1e-12
    assert abs(float(z["norm2"]) - 1) < 1e-12


def test_schedule_model_single_process(oracle):
    """The same schedule applied to an unsharded state with numpy index-bit swaps."""
    from hybridq_b200.dist import plan_sharded
    from hybridq_b200.circuits import sharded_circuit, to_positions
    rng = np.random.default_rng(0)
    for n, g in ((10, 1), (11, 2), (12, 3), (14, 2), (15, 3)):
        nl = n - g
        lowered, _ = to_positions(sharded_circuit(n, g, depth=6, frac_global=0.3, seed=n), qubits=list(range(n)))
        psi = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
        ops, stats, where = plan_sharded(lowered, n, g)
        st = psi.copy()
        done = []
        for op in ops:
            if op.kind == "local":
                done += op.gate_ids
                for U, pos in op.gates:
                    assert all(p < nl for p in pos)          # never touches a rank bit
                    st = oracle.numpy_apply_U(st, U, pos)
            elif op.kind == "permute":
                st = oracle.numpy_swap(st, list(op.perm) + list(range(nl, n)))
            else:
                perm = list(range(n))                      # rank bit gb <-> local bit lp, in place
                for gb, lp in zip(op.gbits, op.lpos):
                    assert lp < nl
                    perm[lp], perm[nl + gb] = nl + gb, lp
                st = oracle.numpy_swap(st, perm)
        assert sorted(done) == list(range(len(lowered)))
        assert where == list(range(n))
        assert np.abs(st - oracle.evolve_numpy(psi, lowered)).max() < 1e-11
