"""Lone dense complex64 k = 4 / 5 gates through the library: tcgen05 kernel (hq_umma.cuh) vs the mma.sync tile-kernel
path it replaces, at several target-position patterns.  Prints one JSON line per case.

    python tools/sweep_umma.py [n_qubits]
"""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

import numpy as np
import torch

import hybridq_b200 as hb
from hybridq_b200.circuits import haar_unitary


def time_plan(plan, st, reps=5):
    plan.run(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def drift(n=16, depth=300, k=5):
    """norm - 1 and max-abs error vs complex128 after `depth` random k-qubit unitaries, tcgen05 vs mma.sync path"""
    rng = np.random.default_rng(77)
    gates = [(haar_unitary(2 ** k, rng), sorted(rng.choice(n, size=k, replace=False).tolist())) for _ in range(depth)]
    psi = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    psi /= np.linalg.norm(psi)
    opts = hb.PlanOptions(fuse=0)
    st64 = hb.DeviceState(n, "complex128").upload(psi)
    hb.Plan(gates, n, "complex128", opts).run(st64)
    ref = st64.download()
    plan = hb.Plan(gates, n, "complex64", opts)
    for mode, name in ((1, "tcgen05"), (0, "mma_sync")):
        old = hb.lib.hq_set_umma(mode)
        st = hb.DeviceState(n, "complex64").upload(psi.astype("complex64"))
        plan.run(st)
        out = st.download()
        hb.lib.hq_set_umma(old)
        print(json.dumps({"test": "drift", "path": name, "n": n, "k": k, "depth": depth,
                          "norm_minus_1": float(np.linalg.norm(out.astype("complex128")) - 1.0),
                          "max_abs_err": float(np.abs(out - ref).max()), "max_abs_amp": float(np.abs(ref).max())}), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "drift":
        for k in (4, 5, 6):
            drift(16, 300, k)
            drift(13, 300, k)
        return
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = np.random.default_rng(3)
    st = hb.DeviceState(n, "complex64")
    st.init_random(1)
    state_gb = 2 * 8 * 2 ** n / 1e9
    cases = [[3, 7, 12, 20, n - 1], [5, 6, 7, 8, 9], [0, 7, 12, 20, n - 2], [0, 1, 12, 20, n - 1], [1, 2, 3, 20, n - 1],
             [0, 1, 2, 3, 4], [n - 5, n - 4, n - 3, n - 2, n - 1], [3, 7, 12, 20], [0, 1, 2, 3], [0, 9, 17, n - 1],
             [n - 4, n - 3, n - 2, n - 1]]
    cases += [[3, 7, 12, 20, n - 3, n - 1], [0, 7, 12, 20, n - 3, n - 1], [0, 1, 2, 3, 4, 5], [n - 6, n - 5, n - 4, n - 3, n - 2, n - 1]]
    for k in (5, 5, 5, 6, 6, 4):
        cases.append(sorted(rng.choice(n, size=k, replace=False).tolist()))
    for pos in cases:
        k = len(pos)
        plan = hb.Plan([(haar_unitary(2 ** k, rng), pos)], n, "complex64")
        row = {"test": "lone_gate_c64", "n": n, "k": k, "pos": pos, "umma_passes": plan.n_umma_passes}
        for mode, name in ((1, "tcgen05"), (0, "mma_sync")):
            old = hb.lib.hq_set_umma(mode)
            ms = time_plan(plan, st)
            hb.lib.hq_set_umma(old)
            row[name + "_ms"] = round(ms, 4)
            row[name + "_GBps"] = round(state_gb / (ms * 1e-3), 1)
        row["speedup"] = round(row["mma_sync_ms"] / row["tcgen05_ms"], 3)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
