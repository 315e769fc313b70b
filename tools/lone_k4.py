#!/usr/bin/env python
"""Lone k = 3 / k = 4 complex64 passes (diagnostics): time and, under ncu, the launch details."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(40)
st = hb.DeviceState(n, "complex64").init_random(seed=1)
for k in (3, 4):
    pos = sorted(int(x) for x in rng.permutation(np.arange(1, 13))[:k])
    plan = hb.Plan([(haar_unitary(2 ** k, rng), pos)], n, "complex64")
    plan.run(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        plan.run(st)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"k": k, "pos": pos, "ms": e0.elapsed_time(e1) / reps}), flush=True)
