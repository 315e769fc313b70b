"""Worker for tests/test_dist_gloo.py: world_size ranks over gloo on CPU.  The local engine is
backed by the oracle (TEST ONLY -- this exercises the sharding schedule and the send/recv
pattern of hybridq_b200.dist, not the CUDA kernels)."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


class OracleEngine:
    def __init__(self, n_local, ctype):
        from oracle import oracle as O
        self.O = O
        self.n_local = n_local
        self.ctype = np.dtype(ctype)

    def alloc(self):
        import torch
        return torch.zeros(2 ** self.n_local, dtype=torch.complex64 if self.ctype == np.complex64 else torch.complex128)

    def tensor(self, st):
        return st

    def run_gates(self, st, key, gates):
        psi = st.numpy()
        for U, pos in gates:
            assert all(p < self.n_local for p in pos)
            psi[:] = self.O.numpy_apply_U(psi, np.asarray(U, dtype=self.ctype), pos)
        return 1

    def permute(self, st, key, perm):
        psi = st.numpy()
        psi[:] = self.O.numpy_swap(psi, perm)
        return 1

    def norm2(self, st):
        return float((st.abs() ** 2).sum())

    def scale(self, st, f):
        st.mul_(f)

    def sync(self):
        pass


def main():
    import torch
    import torch.distributed as dist
    from hybridq_b200.dist import ShardedRunner
    from hybridq_b200.circuits import sharded_circuit, matching_circuit, to_positions
    n, ctype, seed, out_path = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), sys.argv[4]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = int(np.log2(world))
    gates = sharded_circuit(n, g, depth=5, frac_global=0.3, seed=seed) if seed % 2 else matching_circuit(n, depth=4, seed=seed)
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    rng = np.random.default_rng(seed)
    psi = (rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)).astype(ctype)
    psi /= np.linalg.norm(psi)
    runner = ShardedRunner(n, lowered, ctype, dist, engine=OracleEngine(n - g, ctype))
    nl = n - g
    runner.a.copy_(torch.from_numpy(psi[rank * 2 ** nl:(rank + 1) * 2 ** nl].copy()))
    runner.step()
    n2 = runner.norm2()
    full = runner.gather()
    if rank == 0:
        np.savez(out_path, out=full, psi=psi, norm2=n2, exchanges=runner.stats["exchanges"],
                 crossing=runner.stats["crossing_gates"])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
