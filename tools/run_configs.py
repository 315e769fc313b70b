#!/usr/bin/env python
"""BASELINE configs that are not the bench line, measured once per round on one GPU:

  config 3   n = 33, complex128 (128 GiB state, in place), k = 1..6 sweep of 20 Haar gates each on
             uniformly random bits; parity evidence at this size = U then U^dagger returns the
             device-generated initial state (<= 1e-12 max-abs) and the norm is preserved.
  config 5   15-qubit density matrix with depolarizing noise = 2^30 superket, complex64: the lowered
             circuit stored in tests/golden/dm15_circuit.npz (made by the reference's dm front-end);
             checks trace(rho) = 1 and hermiticity.

Writes JSON lines to gpurun_out/configs.jsonl.  Diagnostics / evidence, not the bench number."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

ap = argparse.ArgumentParser()
ap.add_argument("--n3", type=int, default=33)
ap.add_argument("--skip3", action="store_true")
ap.add_argument("--skip5", action="store_true")
args = ap.parse_args()

import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import ksweep_circuit, to_positions, GateApply  # noqa: E402

out = open(ROOT / "gpurun_out" / "configs.jsonl", "a")


def emit(rec):
    out.write(json.dumps(rec) + "\n")
    out.flush()
    print(json.dumps(rec), flush=True)


def timed_run(plan, st):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    plan.run(st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


if not args.skip3:
    n, ctype = args.n3, "complex128"
    st = hb.DeviceState(n, ctype).init_random(seed=33)
    probe = st.tensor[:1 << 20].clone()                   # first 2^20 amplitudes for the element-wise check
    bytes_pass = 2.0 * 2 ** n * 16
    for k in range(1, 7):
        gates = ksweep_circuit(n, k, n_gates=20)
        lowered, _ = to_positions(gates, qubits=list(range(n)))
        inverse = [(U.conj().T, p) for U, p in reversed(lowered)]
        for label, opts in (("one pass per gate", hb.PlanOptions(0, -1, 0, 0, 0, 0, -1)), ("fused", None)):
            fwd = hb.Plan(lowered, n, ctype, opts)
            bwd = hb.Plan(inverse, n, ctype, opts)
            ms_f = timed_run(fwd, st)
            n2 = st.norm2()
            ms_b = timed_run(bwd, st)
            err = float((st.tensor[:1 << 20] - probe).abs().max())
            n2b = st.norm2()
            emit({"config": 3, "n": n, "ctype": ctype, "k": k, "mode": label, "gates": fwd.n_gates,
                  "passes": fwd.n_passes, "ms": ms_f, "gate_applies_per_s": fwd.n_gates / ms_f * 1e3,
                  "GBps_per_pass": bytes_pass * fwd.n_passes / ms_f / 1e6, "norm2_after": n2,
                  "roundtrip_max_abs_err_first_2^20": err, "norm2_after_roundtrip": n2b,
                  "ok": bool(err <= 1e-12 and abs(n2 - 1) < 1e-10)})
    del st, probe
    torch.cuda.empty_cache()

if not args.skip5:
    z = np.load(ROOT / "tests" / "golden" / "dm15_circuit.npz")
    n = int(z["n_super"])
    gates = [GateApply(z[f"g{j}_U"], tuple(int(x) for x in z[f"g{j}_q"])) for j in range(int(z["ngates"]))]
    lowered, _ = to_positions(gates, qubits=list(range(n)))
    ctype = "complex64"
    st = hb.DeviceState(n, ctype).init_product("0" * n)
    plan = hb.Plan(lowered, n, ctype)
    plan.run(st)                                         # warm-up + result
    torch.cuda.synchronize()
    rho = st.tensor.view(2 ** (n // 2), 2 ** (n // 2))
    trace = complex(torch.diagonal(rho).sum().item())
    # hermiticity on a corner block (the full transpose would need a second 8 GiB)
    blk = rho[:4096, :4096]
    herm = float((blk - blk.conj().T).abs().max())
    st.init_product("0" * n)
    ms = timed_run(plan, st)
    ks = np.bincount([len(p) for _, p in lowered], minlength=5).tolist()
    emit({"config": 5, "n_super": n, "ctype": ctype, "gates": plan.n_gates, "k_hist": ks,
          "kernel_matrices": plan.n_kernel_gates, "passes": plan.n_passes, "ms": ms,
          "gate_applies_per_s": plan.n_gates / ms * 1e3, "trace_re": trace.real, "trace_im": trace.imag,
          "hermiticity_max_abs_4096_block": herm, "ok": bool(abs(trace - 1) < 1e-4 and herm < 1e-6)})
out.close()
