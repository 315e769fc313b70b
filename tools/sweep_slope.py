#!/usr/bin/env python
"""Marginal cost of one more kernel matrix in a saturated pass (diagnostics; JSON lines).
One pass with G unmerged gates of k qubits on random bits of the low tile window, G = 1..12;
the slope between G = 6 and G = 12 is the per-matrix compute cost once HBM time is hidden."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--n64", type=int, default=30)
ap.add_argument("--n128", type=int, default=29)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402


def timed(plan, st, reps):
    plan.run(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        plan.run(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for ctype, n, window in (("complex64", args.n64, 13), ("complex128", args.n128, 12)):
    st = hb.DeviceState(n, ctype).init_random(seed=1)
    for k in (1, 2, 3, 4):
        for mma in (0, 2):
            if k == 1 and mma:
                continue
            rng = np.random.default_rng(10 * k)
            gates = [(haar_unitary(2 ** k, rng), sorted(int(x) for x in rng.permutation(np.arange(1, window))[:k]))
                     for _ in range(12)]
            res = {}
            for G in (1, 2, 4, 6, 8, 12):
                plan = hb.Plan(gates[:G], n, ctype, hb.PlanOptions(0, 1, 1, 0, 0, 0, -1, 1, mma))
                assert plan.n_passes == 1 and plan.n_kernel_gates == G
                res[G] = timed(plan, st, args.reps)
            print(json.dumps({"ctype": ctype, "n": n, "k": k, "mma_min_k": mma, "ms_by_G": res,
                              "slope_ms_per_matrix": (res[12] - res[6]) / 6}), flush=True)
    del st
    torch.cuda.empty_cache()
