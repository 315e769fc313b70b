#!/usr/bin/env python
"""Small launch sequence for `ncu --set full`: single-gate passes (k = 1, 2 on mid/low/high bits)
and the first passes of the fused benchmark circuit.  Diagnostics only."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=28)
ap.add_argument("--ctype", default="complex64")
ap.add_argument("--fused-passes", type=int, default=3)
args = ap.parse_args()

import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary, matching_circuit, to_positions  # noqa: E402

n, ctype = args.n, args.ctype
rng = np.random.default_rng(0)
st = hb.DeviceState(n, ctype).init_random(seed=1)
torch.cuda.synchronize()
cases = [(1, [12]), (2, [5, 11]), (1, [0])]
for k, pos in cases:
    U = haar_unitary(2 ** k, rng)
    hb.Plan([(U, pos)], n, ctype, hb.PlanOptions(0, -1, 0, 0, 0)).run(st)      # tile kernel, one gate
    st.apply(U, pos, direct=True)                                             # direct kernel
gates = matching_circuit(n, depth=20, seed=n)
lowered, _ = to_positions(gates, qubits=list(range(n)))
plan = hb.Plan(lowered, n, ctype)
plan.run(st, 0, args.fused_passes)
torch.cuda.synchronize()
print("launches:", hb.lib.hq_launch_count(), "passes in plan:", plan.n_passes,
      [(plan.pass_info(i)["n_gates"], plan.pass_info(i)["n_kernel_gates"]) for i in range(args.fused_passes)])
