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
e2e          same metric through hybridq_b200.simulate(): pinned host state in, pinned host
             state out, H2D + planning + kernels + D2H inside the timed region
roofline     the tile kernel: algorithmic bytes per launch (one read + one write of the state,
             2 * 2^n * 8 B) / mean launch duration, against MEASURED_PEAKS.json
cpu_baseline the reference's own C++ core (oracle/_ref) on this box's host cores, bounded sample
--impl reference   the reference arm: same circuit and state size through the reference core
"""
from __future__ import annotations

import argparse
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


# ------------------------------------------------------------------------------------------
SCALING = "strong"


def workload(n_gpus: int):
    from hybridq_b200.circuits import matching_circuit, sharded_circuit, to_positions
    g = int(round(np.log2(n_gpus)))
    n = N_BASE + (g if SCALING == "weak" else 0)
    if n_gpus == 1 or SCALING == "strong":
        # strong scaling: the very same circuit and state at every N; at N > 1 the top g qubits are
        # the rank, which in this circuit are touched by ~g/15 of the gates (20 % at N = 8)
        gates = matching_circuit(n, depth=DEPTH, seed=n)
        name = f"{n}-qubit depth-{DEPTH} random matching circuit (Haar 1-/2-qubit gates), {CTYPE}, seed {n}"
        if n_gpus > 1:
            name += f", state sharded over {n_gpus} GPUs by the top {g} qubits"
    else:
        gates = sharded_circuit(n, g, depth=DEPTH, frac_global=0.2, seed=n)
        name = (f"{n}-qubit depth-{DEPTH} random circuit, {CTYPE}, top {g} qubits sharded over {n_gpus} GPUs, "
                f"~20% of gates on a sharded qubit, seed {n}")
    lowered, nq = to_positions(gates, qubits=list(range(n)))
    return n, gates, lowered, name


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


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own compiled core (oracle/_ref), each variant
# in its own subprocess (the wheel binary flips FTZ/DAZ; a -march=native build from another
# host may not run here).
# ------------------------------------------------------------------------------------------
_CHILD = r"""
import json, os, sys, time
sys.path.insert(0, {root!r})
import numpy as np
from oracle import oracle as O
import bench
variant, n, n_sample, steps, warmup = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
bench.N_BASE, bench.SCALING = int(sys.argv[6]), sys.argv[7]
core = O.RefCore(variant)
_, _, lowered, _ = bench.workload(1 if n == bench.N_BASE else 2 ** (n - bench.N_BASE))
gates = [(U.astype(bench.CTYPE), p) for U, p in lowered[:n_sample]]
rng = np.random.default_rng(0)
psi = np.zeros(2 ** n, dtype=bench.CTYPE); psi[0] = 1
times = []
for s in range(warmup + steps):
    t = {{}}
    O.evolve_ref(psi, gates, core, timing=t)
    if s >= warmup:
        times.append(t["gate_loop_s"])
print(json.dumps({{"variant": variant, "times": times, "n_sample": len(gates), "L": core.log2_pack_size}}))
"""


def run_reference_cpu(n: int, n_sample: int, steps: int, warmup: int, timeout: float = 600.0):
    from oracle import oracle as O
    if not (O.REFDIR / "wheel").exists() and not (O.REFDIR / "avx2").exists():
        try:
            O.build(ref=True)
        except Exception:
            pass
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    best = None
    tried = []
    for variant in ("wheel", "avx2"):
        if not O.RefCore.available(variant):
            tried.append(f"{variant}: missing")
            continue
        try:
            r = subprocess.run([sys.executable, "-c", _CHILD.format(root=str(ROOT)), variant, str(n), str(n_sample),
                                str(steps), str(warmup), str(N_BASE), SCALING], env=env, capture_output=True, text=True,
                               timeout=timeout)
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


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU core on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, gates, lowered, name = workload(args.gpus)
    n_cpu = min(n, 30)                     # host RAM bound; larger states are extrapolated by 2^-dn
    n_sample = min(len(lowered), args.ref_sample)
    best, cores, tried = run_reference_cpu(n_cpu, n_sample, args.steps, min(args.warmup, 1))
    if best is None:
        print(json.dumps({"impl": "reference", "unavailable": "no reference core ran: " + "; ".join(tried)}))
        return 0
    variant, t, res = best
    rate = res["n_sample"] / t * (2.0 ** (n_cpu - n))
    sample = (f"first {res['n_sample']} gate-applies of the {len(lowered)}-gate circuit on a 2^{n_cpu} state, "
              f"reference core oracle/_ref/{variant} (pack 2^{res['L']}), gate loop incl. its low-bit swap passes"
              + ("" if n_cpu == n else f"; extrapolated x2^-{n - n_cpu} to n={n}"))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "gate-applies/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": SCALING, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "n_qubits": n, "depth": DEPTH},
            "cpu_baseline": {"value": rate, "unit": "gate-applies/s", "cores": cores, "kind": "reference",
                             "sample": sample, "tried": tried},
            "e2e": {"value": rate, "unit": "gate-applies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-sample", type=int, default=24, help="gate-applies per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--qubits", type=int, default=0, help="override the base number of qubits (diagnostics only)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
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
    state_bytes = (2 ** n) * 8
    hb.lib.hq_launch_count_reset()

    plan_opts = hb.PlanOptions(*[int(x) for x in args.plan_options.split(",")]) if args.plan_options else None
    if world == 1:
        runner = SingleGpuRunner(hb, n, lowered, plan_opts)
    else:
        from hybridq_b200.dist import ShardedRunner
        runner = ShardedRunner(n, lowered, CTYPE, dist)
    runner.init_state(seed=n)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

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
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = hb.lib.hq_launch_count() - launches0
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    clocks = sampler.stop() if rank == 0 else None

    n_gates = runner.n_gates
    value = n_gates * steps / (elapsed_ms / 1e3)
    ms_per_step = elapsed_ms / steps

    # ---- roofline of the dominant kernel (tile kernel): per-launch bytes / mean launch time
    kernel_ms = runner.kernel_time_ms(reps=2)          # CUDA events around the local launches only
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

    # ---- e2e through the public API (host pinned buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = runner.e2e(gates, steps=max(1, min(steps, 2)), barrier=barrier, dist=dist)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(n, 30)
        best, cores, tried = run_reference_cpu(n_cpu, min(len(lowered), args.ref_sample), steps=1, warmup=0)
        if best is not None:
            variant, t, res = best
            rate = res["n_sample"] / t * (2.0 ** (n_cpu - n))
            cpu = {"value": rate, "unit": "gate-applies/s", "cores": cores, "kind": "reference",
                   "sample": f"first {res['n_sample']} gate-applies of the same circuit on a 2^{n_cpu} state, reference "
                             f"C++/OpenMP core oracle/_ref/{variant} (pack 2^{res['L']}), gate loop incl. its low-bit swaps",
                   "tried": tried}
        else:
            cpu = {"value": None, "unit": "gate-applies/s", "cores": cores, "kind": "reference",
                   "sample": "no reference core ran: " + "; ".join(tried)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "gate-applies/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": SCALING,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": name, "n_qubits": n, "depth": DEPTH, "gate_applies_per_step": n_gates,
                           "state_bytes_per_gpu": state_bytes // world,
                           "l2": f"state shard ({state_bytes // world >> 20} MiB per GPU) is larger than the 126 MB L2; "
                                 "no flush needed",
                           "parallelism": runner.describe()},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


class SingleGpuRunner:
    def __init__(self, hb, n, lowered, plan_opts=None):
        self.hb = hb
        self.n = n
        self.lowered = lowered
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
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            self.plan.run(self.state)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps

    def extra_roofline(self, peak):
        """Lone 1-/2-qubit gate launches (the north star's >= 70 % HBM target) and the FP32 rate of
        the fused passes, measured live with CUDA events."""
        import torch
        from hybridq_b200.circuits import haar_unitary
        rng = np.random.default_rng(1)
        out = {}
        bytes_pass = 2.0 * (2 ** self.n) * 8
        single = {}
        for name, k, pos in (("k1_bit12", 1, [12]), ("k1_bit0", 1, [0]), ("k1_top", 1, [self.n - 1]),
                             ("k2_bits5_11", 2, [5, 11]), ("k2_bits0_top", 2, [0, self.n - 1])):
            U = haar_unitary(2 ** k, rng)
            plan = self.hb.Plan([(U, pos)], self.n, CTYPE)
            for _ in range(3):
                plan.run(self.state)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                plan.run(self.state)
            e1.record()
            torch.cuda.synchronize()
            single[name] = bytes_pass / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9
        worst = min(single.values())
        out["single_gate"] = {"kernel": "hq_direct_kernel", "GBps": single, "min_GBps": worst,
                              "min_frac_of_measured_peak": worst / peak, "min_frac_of_8TBs": worst / 8000.0}
        # lone k = 3..5 gates on random bits: the tensor-core path of the tile kernel (3xTF32 mma.sync)
        tensor = {}
        for k in (3, 4, 5):
            pos = sorted(int(x) for x in rng.permutation(self.n)[:k])
            plan = self.hb.Plan([(haar_unitary(2 ** k, rng), pos)], self.n, CTYPE)
            for _ in range(2):
                plan.run(self.state)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                plan.run(self.state)
            e1.record()
            torch.cuda.synchronize()
            tensor[f"k{k}_random_bits"] = bytes_pass / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e9
        out["tensor_core_gates"] = {"kernel": "hq_tile_kernel, mma.sync 3xTF32 gate path", "GBps": tensor,
                                    "tflops_3xtf32": {f"k{k}": 3 * 8.0 * 2 ** k * 2 ** self.n /
                                                      (bytes_pass / (tensor[f"k{k}_random_bits"] * 1e9)) / 1e12
                                                      for k in (3, 4, 5)}}
        kms = self.kernel_time_ms(reps=1)
        out["fused_fp32_tflops"] = self.plan.flops / (kms * 1e-3) / 1e12
        out["fp32_tflops_nominal_peak"] = 148 * 128 * 2 * 1.965e9 / 1e12
        out["kernel_matrices_per_step"] = self.plan.n_kernel_gates
        return out

    def e2e(self, gates, steps, barrier, dist):
        import torch
        hb = self.hb
        n = self.n
        del self.state
        torch.cuda.empty_cache()
        host_in = torch.empty(2 ** n, dtype=torch.complex64, pin_memory=True)
        host_out = torch.empty(2 ** n, dtype=torch.complex64, pin_memory=True)
        host_in.zero_()
        host_in[0] = 1
        a_in = host_in.numpy().reshape((2,) * n)
        a_out = host_out.numpy()
        hb.simulate(gates, initial_state=a_in, complex_type=CTYPE, out=a_out)          # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            hb.simulate(gates, initial_state=a_in, complex_type=CTYPE, out=a_out)
        barrier()
        dt = time.perf_counter() - t0
        checksum = float(np.abs(a_out[:1024]).sum())
        return {"value": self.n_gates * steps / dt, "unit": "gate-applies/s", "h2d_bytes_per_step": (2 ** n) * 8,
                "d2h_bytes_per_step": (2 ** n) * 8, "ms_per_step": 1e3 * dt / steps, "steps": steps,
                "api": "hybridq_b200.simulate(circuit, initial_state=<pinned host array>, out=<pinned host array>)",
                "result_checksum": checksum}


if __name__ == "__main__":
    sys.exit(main())
