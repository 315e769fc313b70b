#!/usr/bin/env python
"""Launch sequence for `ncu --set full` (round 2): complex64 n = 28, one saturated pass each of
  (1) 8 unmerged k = 2 gates on the constant-bank FFMA2 slots (hq_tile_kernel<float,0,1>),
  (2) 4 unmerged k = 3 gates on the FFMA2 slots (hq_tile_kernel<float,1,1>),
  (3) the same two passes on the ring kernel, and (4) one lone k = 3 gate on the direct kernel.  Diagnostics only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402

n, ctype = 28, "complex64"
st = hb.DeviceState(n, ctype).init_random(seed=1)
torch.cuda.synchronize()
bits = [5, 6, 7, 8, 9, 10, 11, 12]
rng = np.random.default_rng(3)
plans = []
for k, m in ((2, 8), (3, 4)):
    gates = [(haar_unitary(2 ** k, rng), [bits[(k * j + i) % len(bits)] for i in range(k)]) for j in range(m)]
    plans.append(hb.Plan(gates, n, ctype, hb.PlanOptions(merge_max_k=0, mma_min_k=0)))
for mode in (0, 1):
    hb.lib.hq_set_ring(mode)
    for p in plans:
        p.run(st)
        torch.cuda.synchronize()
        print("ring" if mode else "tile", "passes", p.n_passes, "matrices", p.n_kernel_gates)
hb.lib.hq_set_ring(-1)
hb.Plan([(haar_unitary(8, rng), [3, 9, 20])], n, ctype).run(st)
torch.cuda.synchronize()
