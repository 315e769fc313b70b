#!/usr/bin/env python
"""Occupancy experiment (diagnostics): the same fused circuit with library builds that compile the
k <= 2 complex64 tile kernel for 3 / 4 / 5 resident CTAs per SM.  Run once per library via
HYBRIDQ_B200_LIB."""
import json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import hybridq_b200 as hb
from hybridq_b200.circuits import matching_circuit, to_positions
n, ctype = 30, "complex64"
lowered, _ = to_positions(matching_circuit(n, depth=20, seed=n), qubits=list(range(n)))
st = hb.DeviceState(n, ctype).init_random(seed=1)
for T in (11, 12, 13):
    for nbuf in (1, 2):
        hb.lib.hq_set_tuning(nbuf, 0, -1)
        plan = hb.Plan(lowered, n, ctype, hb.PlanOptions(T, 5, 1, 0, 0, 2, -1, 1))
        for _ in range(2):
            plan.run(st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            plan.run(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"lib": os.path.basename(os.environ.get("HYBRIDQ_B200_LIB", "default")), "T": T, "nbuf": nbuf,
                          "passes": plan.n_passes, "ms": ms, "gate_applies_per_s": plan.n_gates / ms * 1e3}), flush=True)
