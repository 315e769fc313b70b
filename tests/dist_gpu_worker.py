"""torchrun worker of tests/test_gpu_scale.py::test_sharded_two_gpus_vs_one_gpu_n26 (one rank per GPU, NCCL).

Evolves the bench circuit (matching_circuit(n, 20, seed=n)) on a state sharded over the ranks and, on rank 0,
on one GPU from the same initial amplitudes; the two final vectors must agree amplitude for amplitude
(max-abs <= 1e-6 complex64: the sharded schedule fuses gates differently, so bit-equality is not expected).
Also exercises hybridq_b200.simulate(..., shard=True) end to end.
"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist
    import hybridq_b200 as hb
    from hybridq_b200.circuits import matching_circuit, to_positions
    from hybridq_b200.dist import ShardedRunner

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    ctype = "complex64"
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    rank, world = dist.get_rank(), dist.get_world_size()

    gates = matching_circuit(n, depth=20, seed=n)
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    runner = ShardedRunner(n, lowered, ctype, dist)
    runner.init_state(seed=n)
    psi0 = runner.gather()                         # full initial state on every rank (canonical order)
    runner.step()
    out = runner.gather()
    ok = True
    if rank == 0:
        st = hb.DeviceState(n, ctype).upload(psi0)
        hb.Plan(lowered, n, ctype).run(st)
        ref = st.download()
        err = float(np.abs(out - ref).max())
        print(f"sharded x{world} vs 1 GPU, n={n}: max-abs diff {err:.3e}; {runner.describe()}", flush=True)
        ok = err <= 1e-6
        del st
    # sharded checkpoint: dump, clobber, load -> the gathered state is unchanged (bit-exact)
    import tempfile
    ckpt = os.path.join(tempfile.gettempdir(), f"hq_ckpt_test_{os.environ.get('MASTER_PORT', '0')}")
    runner.dump(ckpt, chunk_bytes=1 << 24)
    runner.a.tensor.zero_()
    runner.load(ckpt, chunk_bytes=1 << 24)
    ok_ckpt = bool(np.array_equal(runner.gather(), out))
    ok = ok and ok_ckpt
    for f in (f"{ckpt}.rank{rank}of{world}", f"{ckpt}.rank{rank}of{world}.json"):
        try:
            os.remove(f)
        except OSError:
            pass
    # the public entry point on the same circuit: every rank passes the full initial state, gets its shard back
    shard = hb.simulate(gates, initial_state=psi0.reshape((2,) * n), complex_type=ctype, shard=True)
    nl = n - int(round(np.log2(world)))
    mine = out[rank * 2 ** nl:(rank + 1) * 2 ** nl]
    err2 = float(np.abs(shard.reshape(-1) - mine).max())
    flag = torch.tensor([1 if (ok and ok_ckpt and err2 <= 1e-6) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"simulate(shard=True) vs ShardedRunner: max-abs diff {err2:.3e}", flush=True)
        print("SHARDED_PARITY_OK" if int(flag.item()) == 1 else "SHARDED_PARITY_FAILED", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
