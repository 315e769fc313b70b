"""ms per pass of the tile kernel vs the ring kernel for passes of m unmerged gates (n = 30 complex64 by default):
m k = 2 gates on the constant-bank FFMA2 slots, m k = 3 gates on the tensor-core path, m k = 3 gates on the FMA
path.  One JSON line per (kernel, kind, m).  Run on the GPU box: python tools/sweep_ring.py [n]"""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ctype = sys.argv[2] if len(sys.argv) > 2 else "complex64"
rng = np.random.default_rng(0)
st = hb.DeviceState(n, ctype).init_random(seed=1)
hb.lib.hq_set_tuning(-1, -1, 0)          # no direct kernel: m = 1 goes through the tile / ring kernel too
bits = [5, 6, 7, 8, 9, 10, 11, 12, 14, 17, 20, 23]


_P = [np.eye(2), np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]), np.array([[1, 0], [0, -1]])]
DEPOL2 = sum(((1 - 0.01) if a == b == 0 else 0.01 / 15) * np.kron(np.kron(_P[a], _P[b]), np.kron(_P[a], _P[b]).conj())
             for a in range(4) for b in range(4))


def gates_for(kind, m):
    if kind in ("dr1k4", "k4mma"):
        out = []
        for j in range(m):
            pos = [bits[(4 * j + i) % 8] for i in range(4)]
            out.append((DEPOL2 if kind == "dr1k4" else haar_unitary(16, rng), pos))
        return out
    k = 2 if kind in ("k2", "k2generic") else 3
    out = []
    for j in range(m):
        pos = [bits[(k * j + i) % len(bits)] for i in range(k)]
        out.append((haar_unitary(2 ** k, rng), pos))
    return out


def time_plan(plan, reps=5):
    for _ in range(2):
        plan.run(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        plan.run(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for kind, opts in (("k2", dict(merge_max_k=0)), ("k3mma", dict(merge_max_k=0, mma_min_k=3)),
                   ("k3fast", dict(merge_max_k=0, mma_min_k=0)), ("k3generic", dict(merge_max_k=0, mma_min_k=0, fast_slots=0)),
                   ("dr1k4", dict(merge_max_k=0)), ("k4mma", dict(merge_max_k=0)), ("k2generic", dict(merge_max_k=0, fast_slots=0))):
    for m in (1, 2, 3, 4, 6, 8):
        plan = hb.Plan(gates_for(kind, m), n, ctype, hb.PlanOptions(**opts))
        row = {"n": n, "ctype": ctype, "kind": kind, "m": m, "passes": plan.n_passes, "kernel_gates": plan.n_kernel_gates}
        for name, mode in (("tile_ms", 0), ("ring_ms", 1)):
            hb.lib.hq_set_ring(mode)
            row[name] = round(time_plan(plan) / max(1, plan.n_passes), 4)
        print(json.dumps(row), flush=True)
hb.lib.hq_set_ring(-1)
