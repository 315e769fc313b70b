#!/usr/bin/env python
"""Norm drift of a deep circuit (diagnostics): depth-D layers of Haar k-qubit gates on random bits, complex64,
FMA paths vs tensor-core path; prints |norm^2 - 1| and the max-abs error against a complex128 run."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import haar_unitary  # noqa: E402

n = 24
rng = np.random.default_rng(4)
for k in (2, 3, 4):
    gates = [(haar_unitary(2 ** k, rng), sorted(int(x) for x in rng.permutation(n)[:k])) for _ in range(600)]
    ref = hb.DeviceState(n, "complex128").init_random(seed=3)
    psi0 = ref.download()
    hb.Plan(gates, n, "complex128", hb.PlanOptions(merge_max_k=0)).run(ref)
    want = ref.download()
    for label, opts in (("fma", hb.PlanOptions(merge_max_k=0, mma_min_k=0)), ("mma", hb.PlanOptions(merge_max_k=0, mma_min_k=2))):
        st = hb.DeviceState(n, "complex64").upload(psi0.astype(np.complex64))
        n0 = st.norm2()
        hb.Plan(gates, n, "complex64", opts).run(st)
        got = st.download()
        print(json.dumps({"k": k, "path": label, "gates": len(gates), "norm2_before": n0, "norm2_after": st.norm2(),
                          "max_abs_err_vs_c128": float(np.abs(got - want).max()),
                          "amp_scale": float(np.abs(want).max())}), flush=True)
