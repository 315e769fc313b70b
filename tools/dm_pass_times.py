#!/usr/bin/env python
"""Config 5 (15-qubit density matrix, 2^30 superket) pass by pass: duration of every pass of the default plan with the
kinds of the matrices it applies.  Diagnostics; prints JSON lines."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402
import hybridq_b200 as hb  # noqa: E402

C = bench.load_circuits()
z = np.load(ROOT / "tests" / "golden" / "dm15_circuit.npz")
n = int(z["n_super"])
gates = [C.GateApply(z[f"g{j}_U"], tuple(int(x) for x in z[f"g{j}_q"])) for j in range(int(z["ngates"]))]
lowered, _ = C.to_positions(gates, qubits=list(range(n)))
st = hb.DeviceState(n, "complex64").init_random(seed=1)
plan = hb.Plan(lowered, n, "complex64")
plan.run(st)
torch.cuda.synchronize()
total = 0.0
for p in range(plan.n_passes):
    info = plan.pass_info(p)
    ks = [len(lowered[g][1]) for g in info["gate_ids"]]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.run(st, p, p + 1)
    e0.record()
    for _ in range(3):
        plan.run(st, p, p + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    total += ms
    print(json.dumps({"pass": p, "ms": round(ms, 3), "kernel_matrices": info["n_kernel_gates"], "gate_applies": len(ks),
                      "k4_channels": ks.count(4), "fast_slots": bin(info["fast_mask"]).count("1"),
                      "tile_bits": info["tile_bits"], "n_high": info["n_high"]}), flush=True)
print(json.dumps({"total_ms": round(total, 2), "sparse_rank_one": plan.n_sparse_rank_one, "arithmetic": plan.arithmetic()}))
