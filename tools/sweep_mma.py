#!/usr/bin/env python
"""Tensor-core path vs FMA paths on the real kernels (diagnostics; prints JSON lines).

  1. lone gates k = 2..6 on random bits through the tile kernel, mma on / off, both precisions
  2. the benchmark circuit (depth-20 matching circuit) under different planner options
"""
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
ap.add_argument("--skip-lone", action="store_true")
args = ap.parse_args()

import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary, matching_circuit, to_positions  # noqa: E402


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


rng = np.random.default_rng(3)
for ctype, n in (("complex64", args.n64), ("complex128", args.n128)):
    st = hb.DeviceState(n, ctype).init_random(seed=1)
    bytes_pass = 2.0 * (2 ** n) * (8 if ctype == "complex64" else 16)
    if not args.skip_lone:
        hb.lib.hq_set_tuning(-1, -1, 0)          # lone k <= 2 gates through the tile kernel too
        for k in range(2, 7):
            for label, pos in (("random", sorted(int(x) for x in rng.permutation(n)[:k])),
                               ("low", list(range(k))), ("high", list(range(n - k, n)))):
                U = haar_unitary(2 ** k, rng)
                U2 = haar_unitary(2 ** k, rng)
                for mma in (0, 2):
                    # two unmerged gates on the same bits: one pass, two kernel matrices
                    plan = hb.Plan([(U, pos), (U2, pos)], n, ctype, hb.PlanOptions(0, 1, 1, 0, 0, 0, -1, 1, mma))
                    ms2 = timed(plan, st, args.reps)
                    plan1 = hb.Plan([(U, pos)], n, ctype, hb.PlanOptions(0, 1, 1, 0, 0, 0, -1, 1, mma))
                    ms1 = timed(plan1, st, args.reps)
                    print(json.dumps({"test": "lone", "ctype": ctype, "n": n, "k": k, "bits": label, "pos": pos, "mma_min_k": mma,
                                      "ms_one_gate": ms1, "GBps_one_gate": bytes_pass / ms1 / 1e6,
                                      "ms_two_gates": ms2, "ms_per_extra_gate": ms2 - ms1}), flush=True)
        hb.lib.hq_set_tuning(-1, -1, 1)
    gates = matching_circuit(n, depth=20, seed=n)
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    variants = [("default", None),
                ("mma off", hb.PlanOptions(mma_min_k=0)),
                ("merge 2 analytic (old default)", hb.PlanOptions(merge_max_k=2, merge_pass_cost=12)),
                ("merge 3 table", hb.PlanOptions(merge_max_k=3)),
                ("merge 4 table, mma>=2", hb.PlanOptions(mma_min_k=2)),
                ("merge 3 analytic cost 12", hb.PlanOptions(merge_max_k=3, merge_pass_cost=12)),
                ("merge 3 analytic cost 2", hb.PlanOptions(merge_max_k=3, merge_pass_cost=2)),
                ("merge 4 analytic cost 30", hb.PlanOptions(merge_max_k=4, merge_pass_cost=30))]
    for label, opts in variants:
        plan = hb.Plan(lowered, n, ctype, opts)
        ms = timed(plan, st, args.reps)
        print(json.dumps({"test": "circuit", "ctype": ctype, "n": n, "variant": label, "ms_per_step": ms,
                          "gate_applies_per_s": plan.n_gates / ms * 1e3, "passes": plan.n_passes,
                          "kernel_matrices": plan.n_kernel_gates, "tflops": plan.flops / ms / 1e9}), flush=True)
    del st
    torch.cuda.empty_cache()
