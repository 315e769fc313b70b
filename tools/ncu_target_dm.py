#!/usr/bin/env python
"""ncu target: passes 15 and 9 of the config-5 plan (15-qubit density matrix, 2^30 superket) at full size."""
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
for p in (15, 9):
    plan.run(st, p, p + 1)
    torch.cuda.synchronize()
