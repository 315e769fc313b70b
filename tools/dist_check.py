#!/usr/bin/env python
"""Multi-GPU parity + timing check (launch with torch.distributed.run, one rank per GPU):
the sharded evolution (hybridq_b200.dist) must reproduce the single-GPU evolution of the same
circuit and initial state.  Writes one JSON line per case to gpurun_out/dist_check.jsonl (rank 0)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", type=int, nargs="+", default=[24, 28])
ap.add_argument("--ctypes", nargs="+", default=["complex64", "complex128"])
ap.add_argument("--timing-size", type=int, default=0, help="also time a larger sharded run (no gather)")
args = ap.parse_args()

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import sharded_circuit, to_positions  # noqa: E402
from hybridq_b200.dist import ShardedRunner  # noqa: E402

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
g = int(np.log2(world))
out_path = ROOT / "gpurun_out" / "dist_check.jsonl"
out_path.parent.mkdir(exist_ok=True)
lines = []

for n in args.sizes:
    for ctype in args.ctypes:
        gates = sharded_circuit(n, g, depth=8, frac_global=0.25, seed=n)
        lowered, _ = to_positions(gates, qubits=list(range(n)))
        runner = ShardedRunner(n, lowered, ctype, dist)
        runner.init_state(seed=7)
        psi0 = runner.gather() if rank == 0 or True else None      # all_gather is collective
        runner.step()
        torch.cuda.synchronize()
        n2 = runner.norm2()
        full = runner.gather()
        if rank == 0:
            st = hb.DeviceState(n, ctype).upload(psi0)
            hb.Plan(lowered, n, ctype).run(st)
            ref = st.download()
            err = float(np.abs(full - ref).max())
            tol = 1e-6 if ctype == "complex64" else 1e-12
            rec = {"n": n, "ctype": ctype, "world": world, "max_abs_err_vs_1gpu": err, "tol": tol,
                   "ok": bool(err <= tol), "norm2": n2, "stats": runner.stats}
            print(json.dumps(rec), flush=True)
            lines.append(rec)
            del st
        del runner
        torch.cuda.empty_cache()
        dist.barrier()

# the public entry point on a sharded state: every rank calls simulate() with the same arguments
from hybridq_b200.circuits import matching_circuit  # noqa: E402
n = 22
gates = matching_circuit(n, depth=6, seed=5)
shard, info = hb.simulate(gates, initial_state="+-01" * 5 + "0+", complex_type="complex128", return_info=True)
full_parts = [torch.empty(shard.size, dtype=torch.complex128, device="cuda") for _ in range(world)]
dist.all_gather(full_parts, torch.from_numpy(shard.reshape(-1).copy()).cuda())
if rank == 0:
    ref = hb.simulate(gates, initial_state="+-01" * 5 + "0+", complex_type="complex128", shard=False).reshape(-1)
    err = float(np.abs(torch.cat(full_parts).cpu().numpy() - ref).max())
    rec = {"simulate_sharded_n": n, "world": world, "max_abs_err_vs_1gpu": err, "ok": bool(err <= 1e-12),
           "info": {k: v for k, v in info.items() if k in ("n_passes", "n_gate_applies", "shard", "exchange stats")}}
    print(json.dumps(rec), flush=True)
    lines.append(rec)
dist.barrier()

# Projection and Measure on the sharded state (targets on rank bits and local bits), every rank seeds numpy alike
from hybridq_b200.circuits import ProjectionApply, MeasureApply  # noqa: E402
n = 22
g1 = matching_circuit(n, depth=3, seed=6)
circ = g1 + [ProjectionApply((0, 7), "10")] + matching_circuit(n, depth=2, seed=7) + [MeasureApply((1, 0, 12))] + \
    matching_circuit(n, depth=2, seed=8)
np.random.seed(99)
shard = hb.simulate(circ, initial_state="+" * n, complex_type="complex128")
full_parts = [torch.empty(shard.size, dtype=torch.complex128, device="cuda") for _ in range(world)]
dist.all_gather(full_parts, torch.from_numpy(shard.reshape(-1).copy()).cuda())
if rank == 0:
    np.random.seed(99)
    ref = hb.simulate(circ, initial_state="+" * n, complex_type="complex128", shard=False).reshape(-1)
    got = torch.cat(full_parts).cpu().numpy()
    err = float(np.abs(got - ref).max())
    rec = {"simulate_sharded_functional_n": n, "world": world, "max_abs_err_vs_1gpu": err, "ok": bool(err <= 1e-12),
           "norm2": float(np.vdot(got, got).real)}
    print(json.dumps(rec), flush=True)
    lines.append(rec)
dist.barrier()

if args.timing_size:
    n = args.timing_size
    gates = sharded_circuit(n, g, depth=20, frac_global=0.2, seed=n)
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    runner = ShardedRunner(n, lowered, "complex64", dist)
    runner.init_state(seed=n)
    for _ in range(2):
        runner.step()
    torch.cuda.synchronize()
    dist.barrier()
    runner.exchange_ms = 0.0
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        runner.step(time_exchange=True)
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t0) / reps
    kern = runner.kernel_time_ms(reps=1)
    if rank == 0:
        rec = {"timing_n": n, "world": world, "ms_per_step": 1e3 * dt, "gate_applies_per_s": runner.n_gates / dt,
               "exchange_ms_per_step": runner.exchange_ms / reps, "local_kernel_ms_per_step": kern,
               "norm2": None, "describe": runner.describe()}
        print(json.dumps(rec), flush=True)
        lines.append(rec)
    n2 = runner.norm2()
    if rank == 0:
        lines[-1]["norm2"] = n2

if rank == 0:
    with open(out_path, "a") as f:
        for rec in lines:
            f.write(json.dumps(rec) + "\n")
dist.destroy_process_group()
