#!/usr/bin/env python
"""Launch sequence for `ncu --set full` on the tcgen05 lone-gate kernel (hq_umma.cuh), through the library:
complex64 n = 28, one dense k = 5 gate and one dense k = 4 gate at spread-out targets, one k = 5 gate on the five lowest
bits, one k = 6 gate, then the same k = 5 gate on the mma.sync tile-kernel path it replaces.  Diagnostics only."""
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
rng = np.random.default_rng(3)
plans = [hb.Plan([(haar_unitary(32, rng), [3, 7, 12, 20, 25])], n, ctype),
         hb.Plan([(haar_unitary(16, rng), [3, 7, 12, 20])], n, ctype),
         hb.Plan([(haar_unitary(32, rng), [0, 1, 2, 3, 4])], n, ctype),
         hb.Plan([(haar_unitary(64, rng), [3, 7, 12, 20, 25, 27])], n, ctype)]
for rep in range(2):          # the first launch of each is the warm-up
    for p in plans:
        assert p.n_umma_passes == 1
        p.run(st)
        torch.cuda.synchronize()
hb.lib.hq_set_umma(0)
for rep in range(2):
    plans[0].run(st)
    torch.cuda.synchronize()
hb.lib.hq_set_umma(1)
print("tcgen05 launches", hb.lib.hq_umma_launch_count())
