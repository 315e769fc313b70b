#!/usr/bin/env python
"""Launch sequence for `ncu --set full`: saturated passes (8 unmerged gates on random low bits) --
complex64 k=2 FFMA2 fast slots, k=2 / k=3 / k=4 tensor-core; complex128 k=2 / k=3 tensor-core.  Diagnostics only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402

for ctype, n, window, cases in (("complex64", 28, 13, ((2, 0), (2, 2), (3, 2), (4, 2))), ("complex128", 27, 12, ((2, 2), (3, 2)))):
    st = hb.DeviceState(n, ctype).init_random(seed=1)
    torch.cuda.synchronize()
    for k, mma in cases:
        rng = np.random.default_rng(10 * k)
        gates = [(haar_unitary(2 ** k, rng), sorted(int(x) for x in rng.permutation(np.arange(1, window))[:k])) for _ in range(8)]
        plan = hb.Plan(gates, n, ctype, hb.PlanOptions(0, 1, 1, 0, 0, 0, -1, 1, mma))
        plan.run(st)
        torch.cuda.synchronize()
        print(ctype, "k", k, "mma", mma, "passes", plan.n_passes, "matrices", plan.n_kernel_gates)
    del st
    torch.cuda.empty_cache()
