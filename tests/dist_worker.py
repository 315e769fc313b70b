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

    def marginal(self, st, pos):
        psi = st.numpy()
        idx = np.arange(psi.size)
        s = np.zeros(psi.size, dtype=np.int64)
        for j, p in enumerate(pos):
            s |= ((idx >> int(p)) & 1) << j
        out = np.zeros((2 ** len(pos), 2))
        out[:, 0] = np.bincount(s, weights=psi.real.astype(np.float64) ** 2, minlength=2 ** len(pos))
        out[:, 1] = np.bincount(s, weights=psi.imag.astype(np.float64) ** 2, minlength=2 ** len(pos))
        return out

    def project(self, st, pos, outcome, scale_re, scale_im):
        psi = st.numpy()
        idx = np.arange(psi.size)
        keep = np.ones(psi.size, dtype=bool)
        for j, p in enumerate(pos):
            keep &= ((idx >> int(p)) & 1) == ((outcome >> j) & 1)
        psi[:] = np.where(keep, psi.real * scale_re + 1j * psi.imag * scale_im, 0)

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
    functional = len(sys.argv) > 5 and sys.argv[5] == "functional"
    nl = n - g
    if functional:
        # three gate segments separated by a Projection and a Measure that touch rank bits and local bits
        from hybridq_b200.simulate import _apply_projection, _apply_measure
        from hybridq_b200.circuits import ProjectionApply, MeasureApply
        qmap = {q: n - 1 - q for q in range(n)}
        cut = [len(lowered) // 3, 2 * len(lowered) // 3]
        runner = ShardedRunner(n, lowered[:cut[0]], ctype, dist, engine=OracleEngine(n - g, ctype))
        runner.a.copy_(torch.from_numpy(psi[rank * 2 ** nl:(rank + 1) * 2 ** nl].copy()))
        runner.step()
        _apply_projection(ProjectionApply((0, n - 2), "10"), runner, qmap)
        runner.replan(lowered[cut[0]:cut[1]])
        runner.step()
        np.random.seed(seed)
        _apply_measure(MeasureApply((n - 1, 1, 4)), runner, qmap)
        runner.replan(lowered[cut[1]:])
        runner.step()
    else:
        runner = ShardedRunner(n, lowered, ctype, dist, engine=OracleEngine(n - g, ctype))
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
