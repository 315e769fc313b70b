#!/usr/bin/env python
"""bench.py -- gate-applies/s of the state-vector evolution hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A STEP is one evolution of the whole synthetic circuit (seeded depth-20
random matching circuit of Haar 1-/2-qubit gates) over the state.  N = 1 runs BASELINE
config[1]: n = 30, complex64 (8 GiB state).  N > 1 is STRONG scaling by default: the same
circuit and state, sharded over the ranks by the top log2 N qubits (which ~log2(N)/15 of the
gates touch: 20 % at N = 8, config[3]'s crossing fraction).  `--scaling weak` instead grows the
state with N (n = 30 + log2 N, config[3]'s circuit generator) at a fixed 8 GiB shard per GPU.

value        gate-applies/s, state resident in HBM, timed with CUDA events (max over ranks)
e2e          same metric through hybridq_b200.simulate() at every N: pinned host state (shard) in,
             pinned host state (shard) out, H2D + kernels + exchanges + D2H inside the timed region
roofline     the tile kernel: algorithmic bytes per launch (one read + one write of the state,
             2 * 2^n * 8 B) / mean launch duration, against MEASURED_PEAKS.json
cpu_baseline the reference's own C++ core (oracle/_ref) on this box's host cores, bounded sample
configs      (N = 1, after the timed region) BASELINE configs 3 and 5: n = 33 complex128 k = 1..6 sweep, 15-qubit
             density matrix; (N = 8) config 4: n = 36 complex64 weak scaling with crossing fraction, exposed
             exchange time and the 1-GPU n = 33 denominator
--impl reference   the reference arm: same circuit and state size through the reference core
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "gate-applies/s"
N_BASE = 30
DEPTH = 20
CTYPE = "complex64"
SCALING = "strong"


def load_circuits():
    """hybridq_b200/circuits.py loaded BY FILE PATH: the reference arm must not import the hybridq_b200 package
    (its __init__ dlopens libhybridq_b200.so, which has no business in the reference process)."""
    name = "_hq_bench_circuits"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, ROOT / "hybridq_b200" / "circuits.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------
def workload(n_gpus: int, n_base: int | None = None, scaling: str | None = None):
    C = load_circuits()
    scaling = scaling or SCALING
    g = int(round(np.log2(n_gpus)))
    n = (n_base or N_BASE) + (g if scaling == "weak" else 0)
    if n_gpus == 1 or scaling == "strong":
        # strong scaling: the very same circuit and state at every N; at N > 1 the top g qubits are
        # the rank, which in this circuit are touched by ~g/15 of the gates (20 % at N = 8)
        gates = C.matching_circuit(n, depth=DEPTH, seed=n)
        name = f"{n}-qubit depth-{DEPTH} random matching circuit (Haar 1-/2-qubit gates), {CTYPE}, seed {n}"
        if n_gpus > 1:
            name += f", state sharded over {n_gpus} GPUs by the top {g} qubits"
    else:
        gates = C.sharded_circuit(n, g, depth=DEPTH, frac_global=0.2, seed=n)
        name = (f"{n}-qubit depth-{DEPTH} random circuit, {CTYPE}, top {g} qubits sharded over {n_gpus} GPUs, "
                f"~20% of gates on a sharded qubit, seed {n}")
    lowered, nq = C.to_positions(gates, qubits=list(range(n)))
    return n, gates, lowered, name


def config_block(name: str, n: int, n_gates: int, world: int) -> dict:
    """`config` of the JSON line: identical in both arms (the driver compares them)."""
    state_bytes = (2 ** n) * 8
    return {"workload": name, "n_qubits": n, "depth": DEPTH, "gate_applies_per_step": n_gates,
            "state_bytes_per_gpu": state_bytes // world,
            "l2": f"state shard ({state_bytes // world >> 20} MiB per GPU) is larger than the 126 MB L2; no flush needed"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own compiled core (oracle/_ref), each variant
# in its own subprocess (the wheel binary flips FTZ/DAZ; a -march=native build from another
# host may not run here).  The children import oracle/ and bench.py only -- never hybridq_b200.
# ------------------------------------------------------------------------------------------
_CHILD = r"""
import json, os, sys, time
sys.path.insert(0, {root!r})
import numpy as np
from oracle import oracle as O
import bench
variant, n, n_sample, steps, warmup = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
n_base, scaling, n_gpus = int(sys.argv[6]), sys.argv[7], int(sys.argv[8])
core = O.RefCore(variant)
_, _, lowered, _ = bench.workload(n_gpus, n_base, scaling)
idx = bench.sample_indices(len(lowered), n_sample)
gates = [(lowered[i][0].astype(bench.CTYPE), lowered[i][1]) for i in idx]
psi = np.zeros(2 ** n, dtype=bench.CTYPE); psi[0] = 1
times = []
for s in range(warmup + steps):
    t = {{}}
    O.evolve_ref(psi, gates, core, timing=t)
    if s >= warmup:
        times.append(t["gate_loop_s"])
assert "hybridq_b200" not in sys.modules
print(json.dumps({{"variant": variant, "times": times, "n_sample": len(gates), "L": core.log2_pack_size}}))
"""

# The UNMODIFIED reference package (oracle/_ref/pkg, a git-ignored copy of /root/reference/hybridq) running its own
# simulate(optimize='evolution') on the bench circuit at a reduced n: compress=0 and its default compress=4.
_CHILD_SIMULATE = r"""
import json, sys, time
sys.path.insert(0, {root!r})
import numpy as np
import bench
import hybridq.circuit.simulation.simulation as sim
from hybridq.circuit import Circuit
from hybridq.gate import MatrixGate
n = int(sys.argv[1])
_, gates, lowered, _ = bench.workload(1, n, "strong")
circ = Circuit(MatrixGate(g.U, qubits=list(g.qubits)) for g in gates)
out = {{"n_qubits": n, "gate_applies": len(gates), "log2_pack_size": int(sim._log2_pack_size or 0)}}
for tag, kw in (("compress0", dict(simplify=False, compress=0)), ("compress4_default", dict())):
    t0 = time.perf_counter()
    psi, info = sim.simulate(circ, initial_state="0" * n, optimize="evolution", complex_type=bench.CTYPE,
                             return_info=True, max_largest_intermediate=2 ** 32, **kw)
    out[tag] = {{"gate_loop_s": float(info["runtime (s)"]), "wall_s": time.perf_counter() - t0,
                "norm": float(np.linalg.norm(np.asarray(psi).reshape(-1).astype(np.complex128)))}}
assert "hybridq_b200" not in sys.modules
print(json.dumps(out))
"""


def sample_indices(n_gates: int, n_sample: int):
    """Bounded sample of the circuit for one reference step: n_sample gate-applies spread evenly over ALL layers
    (every gate-apply costs the reference one full pass over the state, plus a swap pass when it touches a low
    bit, so an even spread is representative of the whole circuit)."""
    n_sample = max(1, min(n_sample, n_gates))
    return sorted({int(round(i * (n_gates - 1) / max(1, n_sample - 1))) for i in range(n_sample)})


def run_reference_cpu(n: int, n_sample: int, steps: int, warmup: int, n_gpus: int = 1, timeout: float = 1500.0):
    from oracle import oracle as O
    if not (O.REFDIR / "wheel").exists() and not (O.REFDIR / "avx2").exists():
        try:
            O.build(ref=True)
        except Exception:
            pass
    cores = host_cores()
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    best = None
    tried = []
    for variant in ("wheel", "avx2"):
        if not O.RefCore.available(variant):
            tried.append(f"{variant}: missing")
            continue
        # the slower variant only gets a short look
        st, wu = (steps, warmup) if best is None else (1, 0)
        try:
            r = subprocess.run([sys.executable, "-c", _CHILD.format(root=str(ROOT)), variant, str(n), str(n_sample),
                                str(st), str(wu), str(N_BASE), SCALING, str(n_gpus)], env=env, capture_output=True,
                               text=True, timeout=timeout)
            if r.returncode != 0:
                tried.append(f"{variant}: rc={r.returncode}")
                continue
            res = json.loads(r.stdout.strip().splitlines()[-1])
            t = float(np.mean(res["times"]))
            tried.append(f"{variant}: {res['n_sample'] / t:.3f} gate-applies/s")
            if best is None or t < best[1]:
                best = (variant, t, res)
        except Exception as e:   # timeout, bad json
            tried.append(f"{variant}: {type(e).__name__}")
    return best, cores, tried


def run_reference_simulate(n_small: int, n_full: int, timeout: float = 600.0):
    """The reference's own simulate() end to end (its Python host loop over its C++ core), gate loop timed by the
    reference itself (`info['runtime (s)']`), at n_small qubits and extrapolated x 2^-(n_full - n_small)."""
    ref = ROOT / "oracle" / "_ref"
    lib = next((ref / v for v in ("wheel", "avx2") if (ref / v / "hybridq.so").exists()), None)
    if lib is None or not (ref / "pkg" / "hybridq").exists():
        return {"unavailable": "oracle/_ref/pkg or the reference core is missing"}
    cores = host_cores()
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    env["LD_LIBRARY_PATH"] = f"{lib}:" + env.get("LD_LIBRARY_PATH", "")
    env["PYTHONPATH"] = f"{ref / 'pkg'}:{ref / 'stubs'}"
    try:
        r = subprocess.run([sys.executable, "-W", "ignore", "-c", _CHILD_SIMULATE.format(root=str(ROOT)), str(n_small)],
                           env=env, capture_output=True, text=True, timeout=timeout, cwd="/tmp")
        if r.returncode != 0:
            return {"unavailable": f"rc={r.returncode}: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}"}
        res = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"unavailable": type(e).__name__}
    scale = 2.0 ** (n_small - n_full)
    out = {"api": "hybridq.circuit.simulation.simulate(circuit, initial_state='0..0', optimize='evolution') of the "
                  f"unmodified reference package over oracle/_ref/{lib.name}", "n_qubits_run": n_small,
           "extrapolated_to_n": n_full, "cores": cores, "log2_pack_size": res["log2_pack_size"]}
    for tag in ("compress0", "compress4_default"):
        out[tag] = {"gate_applies_per_s": res["gate_applies"] / res[tag]["gate_loop_s"] * scale,
                    "gate_loop_s_at_n_run": res[tag]["gate_loop_s"], "wall_s_at_n_run": res[tag]["wall_s"]}
    return out


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU core on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, gates, lowered, name = workload(args.gpus)
    n_cpu = min(n, 30)                     # host RAM bound; larger states are extrapolated by 2^-dn
    n_sample = min(len(lowered), args.ref_sample)
    best, cores, tried = run_reference_cpu(n_cpu, n_sample, args.steps, args.warmup, args.gpus)
    if best is None:
        print(json.dumps({"impl": "reference", "unavailable": "no reference core ran: " + "; ".join(tried)}))
        return 0
    variant, t, res = best
    rate = res["n_sample"] / t * (2.0 ** (n_cpu - n))
    sample = (f"{res['n_sample']} gate-applies spread evenly over the {len(lowered)}-gate circuit per step, on a "
              f"2^{n_cpu} state, reference core oracle/_ref/{variant} (pack 2^{res['L']}), driven gate by gate like "
              f"the reference's own host loop incl. its low-bit swap passes (oracle.evolve_ref; compress=0)"
              + ("" if n_cpu == n else f"; extrapolated x2^-{n - n_cpu} to n={n}"))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "gate-applies/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t * len(lowered) / res["n_sample"],
            "higher_is_better": True, "scaling": SCALING, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_block(name, n, len(lowered), args.gpus),
            "parallelism": f"{cores} host threads (OpenMP), no GPU",
            "cpu_baseline": {"value": rate, "unit": "gate-applies/s", "cores": cores, "kind": "reference",
                             "sample": sample, "tried": tried},
            "e2e": {"value": rate, "unit": "gate-applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_ref_simulate:
        line["reference_simulate"] = run_reference_simulate(min(n_cpu, 26), n)
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
def cuda_time_ms(fn, reps: int = 1) -> float:
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=24, help="gate-applies per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 3 / 5 (N = 1) and config 4 (N = 8) blocks")
    ap.add_argument("--no-ref-simulate", action="store_true")
    ap.add_argument("--qubits", type=int, default=0, help="override the base number of qubits (diagnostics only)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--tune", default="", help="diagnostics: hq_set_tuning(nbuf,ctas_per_sm,use_direct), e.g. 2,0,-1")
    ap.add_argument("--plan-options", default="", help="diagnostics: comma-separated PlanOptions fields "
                    "(tile_bits,min_run_bits,fuse,max_gates_per_pass,lookahead,merge_max_k,merge_pass_cost,fast_slots,mma_min_k)")
    args = ap.parse_args()
    global N_BASE, SCALING
    SCALING = args.scaling
    if args.qubits:
        N_BASE = args.qubits
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import hybridq_b200 as hb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    n, gates, lowered, name = workload(world)
    steps, warmup = args.steps, max(args.warmup, 3)
    hb.lib.hq_launch_count_reset()
    if args.tune:
        hb.lib.hq_set_tuning(*[int(x) for x in args.tune.split(",")])

    plan_opts = hb.PlanOptions(*[int(x) for x in args.plan_options.split(",")]) if args.plan_options else None
    if world == 1:
        runner = SingleGpuRunner(hb, n, lowered, plan_opts)
    else:
        # the runner simulate(shard=True) itself uses (cached: shard buffers + peer mappings), so that the e2e leg
        # below runs on the same resident buffers
        runner = hb.sharded_runner(n, CTYPE, dist, plan_opts)
        runner.replan(lowered, key="bench", accumulate=False)
    runner.init_state(seed=n)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(warmup):
        runner.step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = hb.lib.hq_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        runner.step()
    ev1.record()
    barrier()
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = hb.lib.hq_launch_count() - launches0
    if dist is not None:
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    clocks = sampler.stop() if rank == 0 else None

    n_gates = runner.n_gates
    value = n_gates * steps / (elapsed_ms / 1e3)
    ms_per_step = elapsed_ms / steps

    # ---- roofline of the dominant kernel (tile kernel): per-launch bytes / mean launch time
    kernel_ms = max_over_ranks(runner.kernel_time_ms(reps=2))     # CUDA events around the local launches only
    n_local = n - int(round(np.log2(world)))
    bytes_per_launch = 2.0 * (2 ** n_local) * 8
    mean_launch_ms = kernel_ms / max(1, runner.local_passes)
    achieved = bytes_per_launch / (mean_launch_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("tile_kernel_dram_bytes_per_launch")
            if traffic is not None:
                traffic = traffic / world      # measured at n = 30 on one GPU; a shard moves 1/N of it
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "hq_tile_kernel (fused pass)", "peak_source": peak_src,
                "frac_of_8TBs": achieved / 8000.0, "launches_per_step": runner.local_passes,
                "gate_applies_per_launch": n_gates / max(1, runner.local_passes),
                "mean_launch_ms": mean_launch_ms,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "note": "a fused pass moves the state once (algorithmic bytes above) while applying "
                        "gate_applies_per_launch gates; per gate-apply the effective rate is achieved x that factor"}
    roofline.update(runner.extra_roofline(peak))
    parallelism = runner.describe()
    multi = None
    if world > 1:
        # what the exchanges cost on top of the local passes: step time minus the time of the very same passes
        # without the redirect of their write-back (and without the barriers)
        s = runner.stats
        multi = {"crossing_gates": s["crossing_gates"], "crossing_fraction": s["crossing_gates"] / max(1, s["gates"]),
                 "exchanges_per_step": s["exchanges"], "shards_moved_per_gpu_per_step": s["moved_shard_fraction"],
                 "exchange": ("fused into the write-back of the preceding tile pass: peer stores over NVLink (cudaIpc), "
                              "one 4-byte all-reduce as barrier" if runner.fused else "gather + NCCL send/recv"),
                 "local_passes_ms_per_step": kernel_ms,
                 "exposed_exchange_ms_per_step": max(0.0, ms_per_step - kernel_ms),
                 "nvlink_GBps_per_gpu_per_direction_if_all_exposed":
                     s["moved_shard_fraction"] * (2 ** n_local) * 8 / 1e9 / max(1e-9, (ms_per_step - kernel_ms) * 1e-3)}

    # ---- e2e through the public API (host pinned buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = e2e_through_simulate(hb, runner, gates, n, world, rank, steps=max(1, min(steps, 2)), barrier=barrier,
                                   max_over_ranks=max_over_ranks)

    # release the resident state (collective at N > 1: unmaps the peers' buffers) before anything else allocates
    del runner
    hb.clear_caches()
    torch.cuda.empty_cache()

    def guarded(fn, *a):
        """The extra configurations must never cost the headline line: a failure is reported in place."""
        try:
            return fn(*a)
        except Exception as e:          # noqa: BLE001
            torch.cuda.empty_cache()
            return {"error": f"{type(e).__name__}: {e}"[:300]}

    configs = None
    if not args.no_configs and not args.qubits and SCALING == "strong":
        if world == 1:
            configs = {"config3": guarded(run_config3, hb, local_rank), "config5": guarded(run_config5, hb, local_rank)}
        elif world == 8:
            configs = {"config4": guarded(run_config4, hb, dist, rank, local_rank, barrier, max_over_ranks)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(n, 30)
        best, cores, tried = run_reference_cpu(n_cpu, min(len(lowered), args.ref_sample), steps=1, warmup=0)
        if best is not None:
            variant, t, res = best
            rate = res["n_sample"] / t * (2.0 ** (n_cpu - n))
            cpu = {"value": rate, "unit": "gate-applies/s", "cores": cores, "kind": "reference",
                   "sample": f"{res['n_sample']} gate-applies spread evenly over the same circuit on a 2^{n_cpu} state, "
                             f"reference C++/OpenMP core oracle/_ref/{variant} (pack 2^{res['L']}), gate loop incl. its "
                             "low-bit swaps", "tried": tried}
        else:
            cpu = {"value": None, "unit": "gate-applies/s", "cores": cores, "kind": "reference",
                   "sample": "no reference core ran: " + "; ".join(tried)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "gate-applies/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": SCALING,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_block(name, n, n_gates, world), "parallelism": parallelism,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu}
        if multi is not None:
            line["multi_gpu"] = multi
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


class SingleGpuRunner:
    def __init__(self, hb, n, lowered, plan_opts=None):
        self.hb = hb
        self.n = n
        self.lowered = lowered
        self.plan_opts = plan_opts
        self.plan = hb.Plan(lowered, n, CTYPE, plan_opts)
        self.n_gates = self.plan.n_gates
        self.local_passes = self.plan.n_passes
        self.state = hb.DeviceState(n, CTYPE)

    def describe(self):
        return f"1 GPU, {self.plan.n_passes} fused tile passes for {self.n_gates} gate-applies"

    def init_state(self, seed):
        self.state.init_random(seed=seed)

    def step(self):
        self.plan.run(self.state)

    def kernel_time_ms(self, reps=2):
        return cuda_time_ms(lambda: self.plan.run(self.state), reps)

    def extra_roofline(self, peak):
        """Lone 1-/2-/3-qubit gate launches (the north star's >= 70 % HBM target), lone k = 4..6 gates on the
        tensor cores (tcgen05 kernel and the mma.sync path beside it), which arithmetic the fused passes ran on, and the same circuit with complex64 k = 3 matrices on
        the tensor cores (3xTF32 mma.sync, the round-1 default), all measured live with CUDA events."""
        C = load_circuits()
        hb = self.hb
        rng = np.random.default_rng(1)
        out = {}
        bytes_pass = 2.0 * (2 ** self.n) * 8

        def lone(k, pos, reps):
            plan = hb.Plan([(C.haar_unitary(2 ** k, rng), pos)], self.n, CTYPE)
            for _ in range(2):
                plan.run(self.state)
            return bytes_pass / (cuda_time_ms(lambda: plan.run(self.state), reps) * 1e-3) / 1e9

        single = {name: lone(k, pos, 10) for name, k, pos in (
            ("k1_bit12", 1, [12]), ("k1_bit0", 1, [0]), ("k1_top", 1, [self.n - 1]), ("k2_bits5_11", 2, [5, 11]),
            ("k2_bits0_top", 2, [0, self.n - 1]), ("k3_bits3_9_20", 3, [3, 9, 20]), ("k3_bits0_14_top", 3, [0, 14, self.n - 1]))}
        worst = min(single.values())
        out["single_gate"] = {"kernel": "hq_direct_kernel (k <= 3, no shared memory)", "GBps": single, "min_GBps": worst,
                              "min_frac_of_measured_peak": worst / peak, "min_frac_of_8TBs": worst / 8000.0}
        # lone dense k = 4..6 gates: the tcgen05 / TMEM kernel (hq_umma.cuh, default) and the mma.sync tile-kernel
        # path it replaced, same gates, plus the five lowest bits as the worst case for coalescing
        def lone_plan(plan, reps):
            for _ in range(2):
                plan.run(self.state)
            return bytes_pass / (cuda_time_ms(lambda: plan.run(self.state), reps) * 1e-3) / 1e9

        tensor, tensor_old, frac = {}, {}, {}
        umma_before = hb.lib.hq_umma_launch_count()
        for k in (4, 5, 6):
            for name, pos in ((f"k{k}_random_bits", sorted(int(x) for x in rng.permutation(self.n)[:k])),
                              (f"k{k}_lowest_bits", list(range(k)))):
                plan = hb.Plan([(C.haar_unitary(2 ** k, rng), pos)], self.n, CTYPE)
                tensor[name] = lone_plan(plan, 4)
                frac[name] = tensor[name] / peak
                old = hb.lib.hq_set_umma(0)
                try:
                    tensor_old[name] = lone_plan(plan, 2)
                finally:
                    hb.lib.hq_set_umma(old)
        out["tensor_core_gates"] = {
            "kernel": "hq_umma_gate_kernel: tcgen05.mma kind::tf32 (3xTF32), accumulators in TMEM, one gate per pass "
                      "(complex64 k = 4..6)",
            "GBps": tensor, "frac_of_measured_peak": frac,
            "launches": int(hb.lib.hq_umma_launch_count() - umma_before),
            "mma_sync_path_GBps": tensor_old,
            "tflops_3xtf32": {f"k{k}": 3 * 8.0 * 2 ** k * 2 ** self.n /
                              (bytes_pass / (tensor[f"k{k}_random_bits"] * 1e9)) / 1e12 for k in (4, 5, 6)}}
        kms = self.kernel_time_ms(reps=1)
        out["fused_fp32_tflops"] = self.plan.flops / (kms * 1e-3) / 1e12
        out["fp32_tflops_nominal_peak"] = 148 * 128 * 2 * 1.965e9 / 1e12
        out["kernel_matrices_per_step"] = self.plan.n_kernel_gates
        out["arithmetic"] = {"matrices": self.plan.arithmetic(),
                             "note": "complex64 k <= 3 runs on fp32 FMAs (constant-bank FFMA2 slots), the reference's "
                                     "own arithmetic: no tensor-core rounding in `value`"}
        if self.plan_opts is None:
            alt = hb.Plan(self.lowered, self.n, CTYPE, hb.PlanOptions(mma_min_k=3))
            for _ in range(2):
                alt.run(self.state)
            ms = cuda_time_ms(lambda: alt.run(self.state), 3)
            out["alt_tensor_k3"] = {"what": "same circuit with PlanOptions(mma_min_k=3): merged k = 3 matrices on 3xTF32 "
                                            "mma.sync (round-1 default; ~1e-7 per-gate rounding, norm drift 4e-5 / 600 gates)",
                                    "value": self.n_gates / (ms * 1e-3), "ms_per_step": ms,
                                    "matrices": alt.arithmetic()}
        return out


def e2e_through_simulate(hb, runner, gates, n, world, rank, steps, barrier, max_over_ranks):
    """The metric through the public entry point at every N: hybridq_b200.simulate(circuit, initial_state=<pinned
    host array>, out=<pinned host array>[, shard=True]) -- at N > 1 every rank passes and receives its own shard.
    Timed region: H2D of the state, planning (cached after the first call), kernels, exchanges, D2H."""
    import torch
    n_local = n - int(round(np.log2(world)))
    if world == 1:
        # free the resident state first: simulate() allocates its own (at N > 1 it reuses the cached runner)
        if hasattr(runner, "state"):
            del runner.state
        torch.cuda.empty_cache()
    host_in = torch.empty(2 ** n_local, dtype=torch.complex64, pin_memory=True)
    host_out = torch.empty(2 ** n_local, dtype=torch.complex64, pin_memory=True)
    host_in.zero_()
    if rank == 0:
        host_in[0] = 1
    a_in = host_in.numpy().reshape((2,) * n_local)
    a_out = host_out.numpy()
    kw = dict(initial_state=a_in, complex_type=CTYPE, out=a_out)
    if world > 1:
        kw["shard"] = True
    hb.simulate(gates, **kw)          # warm-up (plans, peer mappings)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        hb.simulate(gates, **kw)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    nbytes = (2 ** n) * 8
    return {"value": len(gates) * steps / dt, "unit": "gate-applies/s", "h2d_bytes_per_step": nbytes,
            "d2h_bytes_per_step": nbytes, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "api": "hybridq_b200.simulate(circuit, initial_state=<pinned host array>, out=<pinned host array>"
                   + (", shard=True): every rank passes / receives its shard" if world > 1 else ")"),
            "result_checksum": float(np.abs(a_out[:1024]).sum())}


# ------------------------------------------------------------------------------------------
# BASELINE configs 3, 5 (one GPU) and 4 (8 GPUs), run after the timed region; each with its own clocks sample
# ------------------------------------------------------------------------------------------
def run_config3(hb, gpu_index):
    """n = 33 complex128 (128 GiB, in place): for k = 1..6, 20 Haar gates on random bits (ksweep_circuit, seeds
    331..336) as lone launches and as one fused plan; U then U^dagger must return the initial amplitudes."""
    import torch
    C = load_circuits()
    n, ctype = 33, "complex128"
    free, _total = torch.cuda.mem_get_info()
    if free < (2 ** n) * 16 + (4 << 30):
        n = 32 if free >= (2 ** 32) * 16 + (4 << 30) else 31
    peak, _ = measured_peak()
    sampler = ClockSampler(gpu_index).start()
    st = hb.DeviceState(n, ctype).init_random(seed=33)
    probe = st.tensor[:1 << 20].clone()
    bytes_pass = 2.0 * 2 ** n * 16
    rows = {}
    worst_err = 0.0
    for k in range(1, 7):
        lowered, _ = C.to_positions(C.ksweep_circuit(n, k, n_gates=20), qubits=list(range(n)))
        inverse = [(U.conj().T, p) for U, p in reversed(lowered)]
        row = {}
        for label, opts in (("lone", hb.PlanOptions(fuse=0)), ("fused", None)):
            fwd, bwd = hb.Plan(lowered, n, ctype, opts), hb.Plan(inverse, n, ctype, opts)
            ms = cuda_time_ms(lambda: fwd.run(st))
            bwd.run(st)
            err = float((st.tensor[:1 << 20] - probe).abs().max())
            worst_err = max(worst_err, err)
            row[label] = {"gate_applies_per_s": fwd.n_gates / ms * 1e3, "passes": fwd.n_passes,
                          "GBps_per_pass": bytes_pass * fwd.n_passes / ms / 1e6,
                          "frac_of_measured_peak_per_pass": bytes_pass * fwd.n_passes / ms / 1e6 / peak,
                          "roundtrip_max_abs_err": err}
        rows[f"k{k}"] = row
    n2 = st.norm2()
    del st, probe
    torch.cuda.empty_cache()
    return {"n_qubits": n, "dtype": "complex128", "state_GiB": (2 ** n) * 16 / 2 ** 30, "gates_per_k": 20,
            "sweep": rows, "roundtrip_max_abs_err": worst_err, "norm2_after": n2,
            "ok": bool(worst_err <= 1e-12 and abs(n2 - 1) < 1e-10), "clocks": sampler.stop()}


def run_config5(hb, gpu_index):
    """15-qubit density matrix with depolarizing noise = 2^30 superket, complex64: the lowered circuit made by the
    reference's dm front-end (tests/golden/dm15_circuit.npz); trace and hermiticity of the result."""
    import torch
    C = load_circuits()
    z = np.load(ROOT / "tests" / "golden" / "dm15_circuit.npz")
    n = int(z["n_super"])
    gates = [C.GateApply(z[f"g{j}_U"], tuple(int(x) for x in z[f"g{j}_q"])) for j in range(int(z["ngates"]))]
    lowered, _ = C.to_positions(gates, qubits=list(range(n)))
    ctype = "complex64"
    sampler = ClockSampler(gpu_index).start()
    st = hb.DeviceState(n, ctype).init_product("0" * n)
    plan = hb.Plan(lowered, n, ctype)
    plan.run(st)
    torch.cuda.synchronize()
    rho = st.tensor.view(2 ** (n // 2), 2 ** (n // 2))
    trace = complex(torch.diagonal(rho).sum().item())
    herm = 0.0
    for a in range(0, 2 ** (n // 2), 4096):                 # block-wise: a full transpose would need a second 8 GiB
        blk, blk_t = rho[a:a + 4096, :4096], rho[:4096, a:a + 4096]
        herm = max(herm, float((blk - blk_t.conj().T).abs().max()))
    ms = []
    for _ in range(3):
        st.init_product("0" * n)
        ms.append(cuda_time_ms(lambda: plan.run(st)))
    res = {"n_qubits_dm": n // 2, "n_super": n, "dtype": ctype, "gate_applies": plan.n_gates,
           "k_hist": np.bincount([len(p) for _, p in lowered], minlength=5).tolist(),
           "kernel_matrices": plan.n_kernel_gates, "arithmetic": plan.arithmetic(), "passes": plan.n_passes,
           "ms_per_step": float(np.median(ms)), "gate_applies_per_s": plan.n_gates / float(np.median(ms)) * 1e3,
           "trace_re": trace.real, "trace_im": trace.imag, "hermiticity_max_abs_first_block_row_and_column": herm,
           "ok": bool(abs(trace - 1) < 1e-4 and herm < 1e-6), "clocks": sampler.stop()}
    del st, rho
    torch.cuda.empty_cache()
    return res


def run_config4(hb, dist, rank, local_rank, barrier, max_over_ranks):
    """n = 36 complex64 on 8 GPUs (64 GiB shard per GPU, top 3 qubits = rank), depth-20 circuit with ~20 % of the
    gate-applies on a sharded qubit; then, on rank 0 alone, the 1-GPU denominator at the same shard size (n = 33)."""
    import torch
    from hybridq_b200.dist import ShardedRunner
    C = load_circuits()
    world = dist.get_world_size()
    g = int(round(np.log2(world)))
    n = 33 + g
    free, _total = torch.cuda.mem_get_info()
    while n > 30 + g and 2 * (2 ** (n - g)) * 8 + (6 << 30) > free:
        n -= 1
    gates = C.sharded_circuit(n, g, depth=DEPTH, frac_global=0.2, seed=n)
    lowered, _ = C.to_positions(gates, qubits=list(range(n)))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    runner = ShardedRunner(n, lowered, CTYPE, dist)
    runner.init_state(seed=n)
    runner.step()
    barrier()
    reps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        runner.step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    local_ms = max_over_ranks(runner.kernel_time_ms(reps=1))
    n2 = runner.norm2()
    s = runner.stats
    res = {"n_qubits": n, "dtype": CTYPE, "n_gpus": world, "shard_GiB_per_gpu": (2 ** (n - g)) * 8 / 2 ** 30,
           "gate_applies": len(lowered), "gate_applies_per_s": len(lowered) / ms * 1e3, "ms_per_step": ms,
           "crossing_gates": s["crossing_gates"], "crossing_fraction": s["crossing_gates"] / len(lowered),
           "exchanges_per_step": s["exchanges"], "shards_moved_per_gpu_per_step": s["moved_shard_fraction"],
           "tile_passes_per_step": runner.local_passes, "local_passes_ms_per_step": local_ms,
           "exposed_exchange_ms_per_step": max(0.0, ms - local_ms), "norm2_after": n2,
           "parallelism": runner.describe()}
    runner.close()
    del runner
    torch.cuda.empty_cache()
    barrier()
    # denominator: one GPU, same shard size, same generator without sharded qubits (rank 0 only)
    if rank == 0:
        n1 = n - g
        lowered1, _ = C.to_positions(C.sharded_circuit(n1, 0, depth=DEPTH, frac_global=0.0, seed=n1), qubits=list(range(n1)))
        st = hb.DeviceState(n1, CTYPE).init_random(seed=n1)
        plan = hb.Plan(lowered1, n1, CTYPE)
        plan.run(st)
        ms1 = cuda_time_ms(lambda: plan.run(st), 2)
        res["one_gpu_same_shard"] = {"n_qubits": n1, "gate_applies": plan.n_gates, "passes": plan.n_passes,
                                     "ms_per_step": ms1, "gate_applies_per_s": plan.n_gates / ms1 * 1e3}
        # weak-scaling speed-up: 8 GPUs process a state 8x larger; per gate-apply the work is 8x that of the
        # 1-GPU run, so the speed-up in amplitude-updates per second is 8 * rate(8 GPUs) / rate(1 GPU)
        res["speedup_vs_one_gpu_same_shard"] = world * res["gate_applies_per_s"] / res["one_gpu_same_shard"]["gate_applies_per_s"]
        res["clocks"] = sampler.stop()
        del st
        torch.cuda.empty_cache()
    barrier()
    return res


if __name__ == "__main__":
    sys.exit(main())
