#!/usr/bin/env python
"""Launch sequence for `ncu --set full` on the tensor-core gate path: one pass each with two unmerged
gates of k = 2..6 on random bits, complex64 (3xTF32 mma.sync) and complex128 (FP64 mma.sync).
Diagnostics only."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--n64", type=int, default=28)
ap.add_argument("--n128", type=int, default=27)
args = ap.parse_args()

import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402

rng = np.random.default_rng(0)
for ctype, n in (("complex64", args.n64), ("complex128", args.n128)):
    st = hb.DeviceState(n, ctype).init_random(seed=1)
    torch.cuda.synchronize()
    for k in range(2, 7):
        pos = sorted(int(x) for x in rng.permutation(n)[:k])
        plan = hb.Plan([(haar_unitary(2 ** k, rng), pos), (haar_unitary(2 ** k, rng), pos)], n, ctype,
                       hb.PlanOptions(0, 1, 1, 0, 0, 0, -1, 1, 2))
        plan.run(st)
        torch.cuda.synchronize()
        print(ctype, "k", k, "pos", pos, "passes", plan.n_passes, "matrices", plan.n_kernel_gates)
    del st
    torch.cuda.empty_cache()
