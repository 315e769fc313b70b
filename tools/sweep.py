#!/usr/bin/env python
"""Kernel sweep on one GPU: single-gate passes (tile kernel variants vs the direct kernel)
over target-bit placements, plus whole-circuit plans over planner options.
Writes JSON lines to gpurun_out/sweep.jsonl.  Diagnostics only -- not a bench number."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def timed(fn, warm=2, reps=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n64", type=int, default=30)
    ap.add_argument("--n128", type=int, default=29)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep.jsonl"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-single", action="store_true")
    args = ap.parse_args()
    import torch
    import hybridq_b200 as hb
    from hybridq_b200.circuits import haar_unitary, matching_circuit, to_positions

    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    fout = open(args.out, "w")
    rng = np.random.default_rng(0)

    def emit(rec):
        fout.write(json.dumps(rec) + "\n")
        fout.flush()
        print(json.dumps(rec), flush=True)

    # plain device copy as the local roofline reference
    for ctype, n in (("complex64", args.n64),):
        a = torch.empty(2 ** n, dtype=torch.complex64, device="cuda")
        b = torch.empty_like(a)
        ms = timed(lambda: b.copy_(a))
        emit({"what": "torch_copy", "n": n, "ms": ms, "GBps": 2 * a.numel() * 8 / ms / 1e6})
        del a, b
    torch.cuda.empty_cache()

    for ctype, n in (("complex64", args.n64), ("complex128", args.n128)):
        st = hb.DeviceState(n, ctype).init_random(seed=1)
        bytes_pass = 2 * (2 ** n) * st.complex_type.itemsize
        placements = {
            1: [[0], [1], [2], [4], [7], [12], [20], [n - 1]],
            2: [[0, 1], [0, n - 1], [2, 3], [5, 11], [13, 22], [n - 2, n - 1]],
            3: [[0, 1, 2], [3, 9, 17], [n - 3, n - 2, n - 1]],
            4: [[0, 1, 2, 3], [4, 10, 18, 25], [n - 4, n - 3, n - 2, n - 1]],
            5: [[2, 8, 14, 20, 26]],
            6: [[1, 6, 11, 16, 21, 26]],
        }
        ks = [] if args.no_single else ([1, 2] if args.quick else [1, 2, 3, 4, 5, 6])
        for k in ks:
            U = haar_unitary(2 ** k, rng)
            for pos in placements[k]:
                # direct kernel
                if k <= 3:
                    ms = timed(lambda: st.apply(U, pos, direct=True))
                    emit({"what": "single_gate", "kernel": "direct", "ctype": ctype, "n": n, "k": k, "pos": pos,
                          "ms": ms, "GBps": bytes_pass / ms / 1e6})
                tiles = [11, 12, 13] if ctype == "complex64" else [10, 11, 12]
                for T in tiles:
                    for nbuf in (1, 2):
                        for cps in (0,):
                            hb.lib.hq_set_tuning(nbuf, cps, 0)
                            try:
                                plan = hb.Plan([(U, pos)], n, ctype, hb.PlanOptions(T, -1, 0, 0, 0))
                                ms = timed(lambda: plan.run(st))
                            except Exception as e:
                                emit({"what": "single_gate", "kernel": "tile", "ctype": ctype, "n": n, "k": k,
                                      "pos": pos, "T": T, "nbuf": nbuf, "ctas_per_sm": cps, "error": str(e)})
                                continue
                            info = plan.pass_info(0)
                            emit({"what": "single_gate", "kernel": "tile", "ctype": ctype, "n": n, "k": k, "pos": pos,
                                  "T": T, "nbuf": nbuf, "ctas_per_sm": cps, "n_high": info["n_high"], "ms": ms,
                                  "GBps": bytes_pass / ms / 1e6})
        hb.lib.hq_set_tuning(0, 0, 1)

        # whole circuits: planner options
        gates = matching_circuit(n, depth=20, seed=n)
        lowered, _ = to_positions(gates, qubits=list(range(n)))
        tiles = [12, 13] if ctype == "complex64" else [11, 12]
        for T in tiles:
            for min_run in (5,):
                for merge in (2, 3):
                    for fast in (1, 0):
                        for nbuf in (0, 1):
                            hb.lib.hq_set_tuning(nbuf, 0, -1)
                            try:
                                plan = hb.Plan(lowered, n, ctype, hb.PlanOptions(T, min_run, 1, 0, 0, merge, -1, fast))
                                ms = timed(lambda: plan.run(st), warm=1, reps=2)
                            except Exception as e:
                                emit({"what": "circuit", "ctype": ctype, "n": n, "T": T, "min_run": min_run,
                                      "nbuf": nbuf, "merge": merge, "fast": fast, "error": str(e)})
                                continue
                            emit({"what": "circuit", "ctype": ctype, "n": n, "T": T, "min_run": min_run, "nbuf": nbuf,
                                  "merge": merge, "fast": fast, "gates": plan.n_gates,
                                  "kernel_gates": plan.n_kernel_gates, "passes": plan.n_passes, "ms": ms,
                                  "gate_applies_per_s": plan.n_gates / ms * 1e3, "ms_per_pass": ms / plan.n_passes,
                                  "tflops": plan.flops / ms / 1e9,
                                  "GBps_per_pass": bytes_pass * plan.n_passes / ms / 1e6})
        hb.lib.hq_set_tuning(0, 0, 1)
        # unfused reference point
        plan = hb.Plan(lowered, n, ctype, hb.PlanOptions(0, -1, 0, 0, 0, 0, -1))
        ms = timed(lambda: plan.run(st), warm=1, reps=1)
        emit({"what": "circuit_unfused", "ctype": ctype, "n": n, "gates": plan.n_gates, "passes": plan.n_passes,
              "ms": ms, "gate_applies_per_s": plan.n_gates / ms * 1e3,
              "GBps_per_pass": bytes_pass * plan.n_passes / ms / 1e6})
        del st
        torch.cuda.empty_cache()
    fout.close()


if __name__ == "__main__":
    main()
