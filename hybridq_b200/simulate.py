"""``simulate(circuit, initial_state, optimize='evolution', ...)`` -- host-side mirror of the
reference entry point for the evolution path
(/root/reference/hybridq/circuit/simulation/simulation.py:59 ``simulate`` and :372
``_simulate_evolution``, hybridq branch :464-678), driving the device-resident C ABI.

Same argument names and meaning, same errors for the same misuse, same return values
(final state as a complex ndarray of shape ``(2,)*n``, optionally ``(state, info)`` with
``info['runtime (s)']`` timing the gate loop only, :519/:666/:678).  What differs is where the
state lives: it is uploaded once, every gate runs as part of a fused tile pass on the GPU and
the result is downloaded once.  No per-gate low-bit permutation (:559-630) ever happens -- the
kernels take any target bit -- and there is no split->complex pass at the end (:669-675).

Gate objects are duck-typed on the part of the reference Gate API the hot loop uses
(``gate.qubits``, ``gate.matrix()``, and ``gate.apply(psi, order)`` for FunctionalGates,
:525-554, :633-637), so reference ``Gate``/``Circuit`` objects work unchanged when
``hybridq`` is importable, and :class:`hybridq_b200.circuits.GateApply` works without it.
When ``hybridq`` is importable its own pre-pass (``flatten``, ``simplify``, ``compress``,
``to_matrix_gate``; circuit/utils.py) is used as is -- it is host-side graph rewriting and
out of scope here.
"""
from __future__ import annotations

import time
from typing import Any, Iterable, Sequence
from warnings import warn

import numpy as np

import ctypes
import hashlib
from collections import OrderedDict

from . import _lib
from ._lib import PlanOptions
from .state import DeviceState, Plan

# Compiled plans are cached by the content of the gate stream: calling simulate() again with the same circuit (a
# parameter scan over initial states, a benchmark loop) costs no planning and no program upload.
_PLAN_CACHE: "OrderedDict[tuple, Plan]" = OrderedDict()
_PLAN_CACHE_MAX = 16
_RUNNERS: dict = {}          # (n_qubits, complex type, world size, plan options) -> ShardedRunner


def _gates_key(gates, n_qubits, complex_type, opts) -> tuple:
    h = hashlib.blake2b(digest_size=16)
    for U, pos in gates:
        h.update(np.ascontiguousarray(U, dtype=np.complex128).tobytes())
        h.update(np.asarray(pos, dtype=np.int64).tobytes())
        h.update(b"|")
    o = tuple(getattr(opts, f) for f, _ in opts._fields_) if opts is not None else None
    return (h.hexdigest(), len(gates), int(n_qubits), str(np.dtype(complex_type)), o)


def _cached_plan(gates, n_qubits, complex_type, opts) -> Plan:
    key = _gates_key(gates, n_qubits, complex_type, opts)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        plan = Plan(gates, n_qubits, complex_type, opts)
        _PLAN_CACHE[key] = plan
        while len(_PLAN_CACHE) > _PLAN_CACHE_MAX:
            _PLAN_CACHE.popitem(last=False)
    else:
        _PLAN_CACHE.move_to_end(key)
    return plan


def clear_caches() -> None:
    """Drop the cached plans and release the cached sharded runners (their shard buffers and peer mappings).
    Collective when a sharded runner exists: every rank must call it."""
    _PLAN_CACHE.clear()
    for r in list(_RUNNERS.values()):
        r.close()
    _RUNNERS.clear()


def sharded_runner(n_qubits, complex_type, dist, plan_options=None):
    """The (cached) :class:`hybridq_b200.dist.ShardedRunner` simulate(shard=True) uses for this state size: two
    shard buffers per GPU, mapped by every peer over NVLink.  Collective."""
    from .dist import ShardedRunner
    o = tuple(getattr(plan_options, f) for f, _ in plan_options._fields_) if plan_options is not None else None
    key = (int(n_qubits), str(np.dtype(complex_type)), dist.get_world_size(), o)
    r = _RUNNERS.get(key)
    if r is None:
        for old in list(_RUNNERS.values()):       # one resident sharded state at a time
            old.close()
        _RUNNERS.clear()
        r = _RUNNERS[key] = ShardedRunner(n_qubits, [], complex_type, dist, plan_options=plan_options)
    return r


def _try_hybridq():
    try:
        import hybridq.circuit as hc           # noqa: F401
        import hybridq.circuit.utils as hu     # noqa: F401
        import hybridq.gate.property as pr     # noqa: F401
        return hc, hu, pr
    except Exception:
        return None


def _sorted_qubits(qubits: Iterable) -> list:
    """Sorted list of heterogeneous qubit labels; mirrors hybridq.utils.sort
    (/root/reference/hybridq/utils/utils.py:283): plain ordering when the labels are
    comparable, otherwise ordered by (type name, value)."""
    qs = list(qubits)
    try:
        return sorted(qs)
    except TypeError:
        return sorted(qs, key=lambda x: (type(x).__name__, str(x)) if not isinstance(x, tuple)
                      else ("tuple", tuple(str(y) for y in x)))


def _is_functional(gate) -> bool:
    """FunctionalGate = has ``apply(psi, order)`` and does not provide a matrix
    (reference: isinstance(gate, pr.FunctionalGate), simulation.py:525)."""
    hq = _try_hybridq()
    if hq is not None and isinstance(gate, hq[2].FunctionalGate):
        return True
    if getattr(gate, "name", None) in ("PROJECTION", "MEASURE") and not callable(getattr(gate, "matrix", None)):
        return True
    prov = getattr(gate, "provides", None)
    if callable(prov):
        try:
            if prov(["qubits", "matrix"]):
                return False
        except Exception:
            pass
    return callable(getattr(gate, "apply", None)) and not callable(getattr(gate, "matrix", None))


def _flatten(circuit) -> list:
    out = []
    for g in circuit:
        if hasattr(g, "qubits") and (callable(getattr(g, "matrix", None)) or callable(getattr(g, "apply", None))
                                     or getattr(g, "name", None) in ("PROJECTION", "MEASURE")):
            out.append(g)
        elif hasattr(g, "__iter__"):
            out.extend(_flatten(g))
        else:
            raise RuntimeError(f"'{g}' not supported")
    return out


def simulate(circuit,
             initial_state: Any = None,
             final_state: Any = None,
             optimize: Any = "evolution",
             backend: Any = "numpy",
             complex_type: Any = "complex64",
             tensor_only: bool = False,
             simplify: Any = True,
             remove_id_gates: bool = True,
             use_mpi: bool | None = None,
             atol: float = 1e-8,
             verbose: bool = False,
             **kwargs):
    """Evolve `initial_state` through `circuit` on the GPU.  See the module docstring.

    Extra keyword arguments (all optional): ``compress`` (int or dict, as in the reference;
    default 0 here because gate fusion happens inside the kernel passes), ``max_largest_intermediate``,
    ``return_info``, ``return_numpy_array`` (False returns the :class:`DeviceState`),
    ``plan_options`` (:class:`PlanOptions`), ``device``, ``out`` (preallocated, e.g. pinned, result array).
    """
    if not (isinstance(optimize, str) and "evolution" in optimize):
        raise NotImplementedError(
            "hybridq_b200 implements optimize='evolution' only; use the reference for tensor-network "
            "contraction.")
    if tensor_only:
        raise ValueError(f"'tensor_only' is not support for optimize={optimize}")
    sub = "-".join(optimize.split("-")[1:]) or "hybridq"
    if sub != "hybridq":
        raise NotImplementedError("only optimize='evolution' / 'evolution-hybridq' run on the GPU core")

    kwargs.setdefault("allow_sampling", False)
    kwargs.setdefault("sampling_seed", None)
    kwargs.setdefault("compress", 0)
    kwargs.setdefault("max_largest_intermediate", None)
    kwargs.setdefault("return_info", False)
    kwargs.setdefault("return_numpy_array", True)
    kwargs.setdefault("plan_options", None)
    kwargs.setdefault("device", None)
    kwargs.setdefault("out", None)           # optional (e.g. pinned) complex array receiving the result
    kwargs.setdefault("shard", False)        # True: shard the state over the ranks of an initialised
                                             # torch.distributed process group (one GPU per rank) -- a collective
                                             # call that returns this rank's shard; 'auto': shard when such a group
                                             # exists and the circuit fits the shards, else run unsharded.  The
                                             # default is the reference's behaviour: the full state on every caller.

    hq = _try_hybridq()
    is_ref_circuit = False
    if hq is not None:
        hc, hu, pr = hq
        try:
            circuit = hc.Circuit(circuit)
            is_ref_circuit = True
        except Exception:
            is_ref_circuit = False

    t_pre = time.perf_counter()
    if is_ref_circuit:
        # the reference's own host pre-pass, used as is (simulation.py:232-305, :436-454)
        circuit = hu.flatten(circuit)
        if kwargs["sampling_seed"] is not None:
            _st = np.random.get_state()
            np.random.seed(int(kwargs["sampling_seed"]))
        circuit = hc.Circuit(g.sample() if isinstance(g, pr.StochasticGate) and kwargs["allow_sampling"] else g
                             for g in circuit)
        if kwargs["sampling_seed"] is not None:
            np.random.set_state(_st)
        qubits = circuit.all_qubits()
        if remove_id_gates:
            circuit = hc.Circuit(g for g in circuit if g.name != "I")
        if simplify:
            circuit = hu.simplify(circuit, remove_id_gates=remove_id_gates, atol=atol, verbose=verbose,
                                  **(simplify if isinstance(simplify, dict) else {}))
        if circuit.all_qubits() != qubits:
            raise ValueError("Active qubits have changed after simplification. Forcing stop.")
        comp = kwargs["compress"]
        max_nq = comp["max_n_qubits"] if isinstance(comp, dict) else comp
        if max_nq:
            groups = hu.compress(circuit, max_nq, verbose=verbose, skip_compression=[pr.FunctionalGate],
                                 **({k: v for k, v in comp.items() if k != "max_n_qubits"}
                                    if isinstance(comp, dict) else {}))
            circuit = hc.Circuit(g for c in (c if any(isinstance(g, pr.FunctionalGate) for g in c)
                                             else [hu.to_matrix_gate(c, complex_type=complex_type)]
                                             for c in groups) for g in c)
        gates = list(circuit)
    else:
        gates = _flatten(circuit)
        qubits = _sorted_qubits({q for g in gates for q in g.qubits})    # before identities go (simulation.py:258, :290)
        if remove_id_gates:
            gates = [g for g in gates if getattr(g, "name", None) != "I"]
    n_qubits = len(qubits)

    dist = _dist_if_sharded(kwargs["shard"])
    n_shard_bits = int(round(np.log2(dist.get_world_size()))) if dist is not None else 0

    # initial / final state checks (simulation.py:261-286, :415-426)
    def _prepare(state):
        if isinstance(state, str):
            if len(state) == 1:
                state = state * n_qubits
            if len(state) != n_qubits:
                raise ValueError("Wrong number of qubits for initial/final state.")
            if set(state).difference("+-01"):
                raise ValueError(f"Symbols {set(state).difference('+-01')} are not allowed.")
            return state
        state = np.asarray(state)
        if any(x != 2 for x in state.shape):
            raise ValueError("Only qubits of dimension 2 are supported.")
        if state.ndim != n_qubits and not (n_shard_bits and state.ndim == n_qubits - n_shard_bits):
            raise ValueError("Wrong number of qubits for initial/final state.")
        return state

    initial_state = None if initial_state is None else _prepare(initial_state)
    if final_state is not None:
        warn(f"'final_state' cannot be specified in optimize='{optimize}'. Ignoring 'final_state'.")
    if initial_state is None:
        raise ValueError("'initial_state' must be specified for optimize='evolution'.")
    if kwargs["max_largest_intermediate"] is not None and 2 ** n_qubits > kwargs["max_largest_intermediate"]:
        raise MemoryError("Memory for the given number of qubits exceeds the 'max_largest_intermediate'.")

    complex_type = np.dtype(complex_type)
    if complex_type not in (np.dtype("complex64"), np.dtype("complex128")):
        warn("optimize=evolution-hybridq only support ['complex64', 'complex128']. Using 'complex64'.")
        complex_type = np.dtype("complex64")
    if n_qubits < 1:
        raise ValueError("empty circuit")

    # qubit -> index bit, first sorted qubit = most significant bit (simulation.py:512-513)
    qmap = {q: n_qubits - 1 - i for i, q in enumerate(qubits)}

    # split the gate stream at FunctionalGates; everything else becomes (U, pos)
    segments: list = []
    cur: list = []
    for g in gates:
        if _is_functional(g):
            if cur:
                segments.append(("gates", cur))
                cur = []
            segments.append(("functional", g))
        elif callable(getattr(g, "matrix", None)) and hasattr(g, "qubits"):
            if getattr(g, "name", None) == "I":
                continue                     # kept only to pin the qubit register (remove_id_gates=False)
            U = np.asarray(g.matrix(), dtype=complex_type, order="C")
            pos = [qmap[q] for q in reversed(tuple(g.qubits))]
            cur.append((U, pos))
        else:
            raise RuntimeError(f"'{g}' not supported")
    if cur:
        segments.append(("gates", cur))
    t_pre = time.perf_counter() - t_pre

    if dist is not None and kwargs["shard"] == "auto":
        g_bits = int(round(np.log2(dist.get_world_size())))
        kmax = max((len(p) for kind, seg in segments if kind == "gates" for _, p in seg), default=0)
        if n_qubits - g_bits < max(kmax, 12):        # too small to be worth (or able to be) sharded
            dist = None
    if dist is not None:
        kwargs["_qmap"] = qmap
        return _simulate_sharded(dist, segments, n_qubits, complex_type, initial_state, kwargs, t_pre)

    t_plan = time.perf_counter()
    opts: PlanOptions | None = kwargs["plan_options"]
    plans = [(kind, _cached_plan(payload, n_qubits, complex_type, opts) if kind == "gates" else payload)
             for kind, payload in segments]
    t_plan = time.perf_counter() - t_plan

    state = DeviceState(n_qubits, complex_type, device=kwargs["device"])
    out = kwargs["out"]

    # Pinned host arrays: fold the upload into the first pass and the download into the last one (hq_plan_run_io:
    # the kernels read / write the host arrays directly over PCIe, overlapping the transfers with the arithmetic
    # of those passes).  Needs one gate segment; anything else takes the copy-in / run / copy-out path below.
    def _pinned(a):
        return (a is not None and isinstance(a, np.ndarray) and a.dtype == complex_type and a.flags.c_contiguous
                and a.size == 2 ** n_qubits and bool(_lib.lib.hq_host_is_pinned(ctypes.c_void_p(a.ctypes.data))))

    single = len(plans) == 1 and plans[0][0] == "gates" and plans[0][1].n_passes > 0
    io_src = initial_state.reshape(-1) if (single and not isinstance(initial_state, str) and _pinned(initial_state.reshape(-1))) else None
    io_dst = out.reshape(-1) if (single and kwargs["return_numpy_array"] and out is not None and _pinned(out.reshape(-1))) else None

    t_up = time.perf_counter()
    if io_src is None:
        if isinstance(initial_state, str):
            state.init_product(initial_state)
            state.sync()
        else:
            state.upload(initial_state.reshape(-1))
    t_up = time.perf_counter() - t_up

    # ---- the gate loop (this is what 'runtime (s)' times, as in the reference) ----
    t0 = time.perf_counter()
    n_passes = 0
    n_gate_applies = 0
    if io_src is not None or io_dst is not None:
        plan = plans[0][1]
        plan.run_io(state, io_src, io_dst)
        n_passes, n_gate_applies = plan.n_passes, plan.n_gates
    else:
        for kind, payload in plans:
            if kind == "gates":
                payload.run(state)
                n_passes += payload.n_passes
                n_gate_applies += payload.n_gates
            else:
                _apply_functional(payload, state, qmap)
    state.sync()
    runtime = time.perf_counter() - t0

    info = {"runtime (s)": runtime, "pre-pass (s)": t_pre, "plan (s)": t_plan, "upload (s)": t_up,
            "n_gate_applies": n_gate_applies, "n_passes": n_passes, "n_qubits": n_qubits,
            # passes on the tcgen05 / TMEM kernel (one dense complex64 4..6-qubit matrix each) and channels that run
            # in the sparse scalar + rank-one form
            "n_tcgen05_passes": sum(p.n_umma_passes for kind, p in plans if kind == "gates"),
            "n_sparse_channels": sum(p.n_sparse_rank_one for kind, p in plans if kind == "gates"),
            "transfers folded into passes": {"upload": io_src is not None, "download": io_dst is not None}}
    if kwargs["return_numpy_array"]:
        t_down = time.perf_counter()
        if io_dst is not None:
            psi = out.reshape((2,) * n_qubits)
        else:
            psi = state.download(out.reshape(-1) if out is not None else None).reshape((2,) * n_qubits)
        info["download (s)"] = time.perf_counter() - t_down
    else:
        psi = state
    return (psi, info) if kwargs["return_info"] else psi


def expectation_value(state, op, qubits_order, complex_type: Any = "complex64", backend: Any = "numpy",
                      verbose: bool = False, **kwargs):
    """<state| op |state> with the evolution on the GPU: mirror of the reference's
    ``expectation_value`` (/root/reference/hybridq/circuit/simulation/simulation.py:1125-1217): same
    arguments, same checks, same result ``sum(simulate(op, state) * conj(state))``.  The evolved
    state never leaves the device: the inner product is a device reduction (``hq_vdot_dev``)."""
    kwargs["remove_id_gates"] = False
    state = np.asarray(state)
    n_qubits = state.ndim
    qubits_order = list(qubits_order)
    if len(qubits_order) != n_qubits:
        raise ValueError("'qubits_order' must have the same number of qubits of 'state'.")
    op = list(_flatten(op))
    used = {q for g in op for q in g.qubits}
    if used.difference(qubits_order):
        raise ValueError("'op' has qubits not included in 'qubits_order'.")
    from .circuits import GateApply
    op = op + [GateApply(np.eye(2), (q,), name="I") for q in set(qubits_order).difference(used)]
    kwargs.pop("return_numpy_array", None)
    kwargs.pop("return_info", None)
    evolved = simulate(op, initial_state=state, optimize="evolution", complex_type=complex_type, backend=backend,
                       verbose=verbose, return_numpy_array=False, shard=False, **kwargs)
    bra = DeviceState(n_qubits, evolved.complex_type, device=evolved.device).upload(state.reshape(-1))
    return np.real_if_close(np.asarray(bra.vdot(evolved), dtype=evolved.complex_type))


def _dist_if_sharded(shard):
    """torch.distributed module if the state is to be sharded over its ranks, else None."""
    if shard is False:
        return None
    try:
        import torch.distributed as dist
    except Exception:
        dist = None
    ok = dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if shard is True and not ok:
        raise RuntimeError("shard=True needs an initialised torch.distributed process group with > 1 rank")
    return dist if ok else None


def _simulate_sharded(dist, segments, n_qubits, complex_type, initial_state, kwargs, t_pre):
    """Multi-GPU evolution (hybridq_b200.dist): every rank calls simulate() with the same circuit.  An array
    initial state is either the FULL state on every rank (each uploads its own slice) or, with ndim =
    n - log2(world), this rank's shard (amplitudes [rank * 2^(n-g), (rank+1) * 2^(n-g)) of the full state).
    Returns this rank's shard of the final state in canonical order.  The runner (shard buffers, peer mappings)
    and the schedules / compiled plans are cached across calls.  The reference has no counterpart
    (simulation.py:379-380)."""
    for kind, payload in segments:
        if kind != "gates" and getattr(payload, "name", None) not in ("PROJECTION", "MEASURE"):
            raise NotImplementedError("only Projection and Measure FunctionalGates are supported on a sharded state")
    qmap = kwargs.pop("_qmap")
    t_plan = time.perf_counter()
    runner = sharded_runner(n_qubits, complex_type, dist, kwargs["plan_options"])
    runner.stats = {}
    runner.n_gates = 0
    keys = [(_gates_key(payload, n_qubits, complex_type, None)[0] if kind == "gates" else None)
            for kind, payload in segments]
    t_plan = time.perf_counter() - t_plan
    nl = runner.n_local
    t_up = time.perf_counter()
    if isinstance(initial_state, str):
        runner.init_product(initial_state)
    else:
        flat = np.asarray(initial_state).reshape(-1)
        if flat.size == 2 ** n_qubits:
            flat = flat[runner.rank * 2 ** nl:(runner.rank + 1) * 2 ** nl]
        runner.load_shard(flat if flat.dtype == complex_type and flat.flags.c_contiguous
                          else np.ascontiguousarray(flat, dtype=complex_type))
    t_up = time.perf_counter() - t_up
    t0 = time.perf_counter()
    for (kind, payload), key in zip(segments, keys):
        if kind == "gates":
            runner.replan(payload, key=key)
            runner.step()
        elif payload.name == "PROJECTION":
            _apply_projection(payload, runner, qmap)
        else:
            _apply_measure(payload, runner, qmap)
    runner.engine.sync()
    runtime = time.perf_counter() - t0
    info = {"runtime (s)": runtime, "pre-pass (s)": t_pre, "plan (s)": t_plan, "upload (s)": t_up,
            "n_gate_applies": runner.n_gates, "n_passes": runner.local_passes, "n_qubits": n_qubits,
            "shard": (runner.rank, runner.world), "n_local_qubits": nl, "exchange stats": runner.stats}
    if kwargs["return_numpy_array"]:
        out = kwargs["out"]
        psi = runner.download_shard(out.reshape(-1) if out is not None else None).reshape((2,) * nl)
    else:
        psi = runner.a
    return (psi, info) if kwargs["return_info"] else psi


_PROJECTION_ATOL = 1e-6      # hybridq/gate/projection.py:31


def _apply_projection(gate, state: DeviceState, qmap: dict, renormalize: bool = True) -> None:
    """ProjectionGate on the device, any number of qubits.  Follows the reference on split planes
    (projection.py:70-116 with _Projection :25-68 on each plane): a plane whose projected norm is <= 1e-6 is
    zeroed as a whole, then the state is renormalised by what is left.  Only the one outcome's re^2 / im^2 sums
    are needed, so the reduction is a conditional sum (no 2^k histogram)."""
    spec = tuple(gate.state)
    if len(spec) != len(tuple(gate.qubits)):
        raise ValueError("'state' is not consistent with 'axes'.")
    if any(x not in ("0", "1", 0, 1) for x in spec):
        raise ValueError("Only projections to the z-basis are supported at the moment.")
    pos = [qmap[q] for q in gate.qubits]
    outcome = sum(int(b) << j for j, b in enumerate(spec))
    if hasattr(state, "broadcast_int"):                 # sharded state (hybridq_b200.dist.ShardedRunner)
        sums = state.marginal(pos)[outcome]
    else:
        mask = sum(1 << p for p in pos)
        value = sum(int(b) << p for p, b in zip(pos, spec))
        sums = state.marginal([], cond_mask=mask, cond_value=value)[0]
    keep = [float(np.sqrt(x) > _PROJECTION_ATOL) for x in sums]
    scale = 1.0
    if renormalize:
        norm = float(np.sqrt(sums[0] * keep[0] + sums[1] * keep[1]))
        if norm != 0:
            scale = 1.0 / norm
    state.project(pos, outcome, keep[0] * scale, keep[1] * scale)


_MEASURE_CHUNK = 20          # outcome bits per reduction when more than 24 qubits are measured at once


def _apply_measure(gate, state: DeviceState, qmap: dict, renormalize: bool = True) -> int:
    """MeasureGate on the device (measure.py:25-75): outcome probabilities by a device reduction, the draw
    with numpy's global generator exactly as the reference does (`np.random.choice(size, p=probs)`, :52 -- the
    probabilities are handed over as they are, in the state's real precision, so an unnormalised state is
    rejected by numpy here as it is there), projection + renormalisation by a second kernel.  Outcome index:
    gate.qubits[0] is the most significant digit (the reference transposes the measured axes to the front in
    gate.qubits order, :39-43).

    Up to 24 qubits the draw is the reference's own (same generator state -> same outcome).  Beyond that the
    2^k probability vector the reference builds on the host does not fit anywhere sensible, so the outcome is
    drawn digit group by digit group (most significant first, `_MEASURE_CHUNK` bits at a time, each from the
    marginal conditioned on the groups already drawn): the same distribution, a different use of the
    generator."""
    qubits = tuple(gate.qubits)
    k = len(qubits)
    pos = [qmap[q] for q in reversed(qubits)]           # outcome bit j <-> qubits[k-1-j]
    ft = np.float32 if state.complex_type == np.complex64 else np.float64
    sharded = hasattr(state, "broadcast_int")
    if k <= 24 or sharded:
        sums = state.marginal(pos)
        probs = sums.sum(axis=1).astype(ft)
        if sharded:                                     # rank 0 draws for everybody
            outcome = state.broadcast_int(int(np.random.choice(2 ** k, p=probs)) if state.rank == 0 else 0)
        else:
            outcome = int(np.random.choice(2 ** k, p=probs))
        p_outcome = float(sums[outcome].sum())
    else:
        outcome, mask, value, p_prev = 0, 0, 0, 1.0
        hi = k
        while hi > 0:
            lo = max(0, hi - _MEASURE_CHUNK)
            chunk = pos[lo:hi]
            sums = state.marginal(chunk, cond_mask=mask, cond_value=value).sum(axis=1)
            probs = sums / sums.sum()
            s = int(np.random.choice(len(probs), p=probs))
            outcome |= s << lo
            for j, p in enumerate(chunk):
                mask |= 1 << p
                value |= ((s >> j) & 1) << p
            p_prev = float(sums[s])
            hi = lo
        p_outcome = p_prev
    scale = 1.0
    if renormalize:
        scale = 1.0 / float(np.sqrt(p_outcome))
    state.project(pos, outcome, scale, scale)
    return outcome


def _apply_functional(gate, state: DeviceState, qmap: dict) -> None:
    """FunctionalGate contract of the reference (simulation.py:525-554): the gate receives
    the state as a real array of shape (2,)+(2,)*n (re/im planes) plus the qubit order and
    returns (new_psi, new_order).  Projection and Measure gates (the reference's own FunctionalGates)
    run on the device; anything else is a D2H/H2D round trip around arbitrary Python."""
    name = getattr(gate, "name", None)
    if name == "PROJECTION" and hasattr(gate, "state"):
        return _apply_projection(gate, state, qmap)
    if name == "MEASURE":
        _apply_measure(gate, state, qmap)
        return None
    n = state.n_qubits
    order = tuple(q for q, _ in sorted(qmap.items(), key=lambda x: x[1])[::-1])
    psi = state.download()
    ft = np.float32 if state.complex_type == np.complex64 else np.float64
    planes = np.empty((2,) + (2,) * n, dtype=ft)
    planes[0] = psi.real.reshape((2,) * n)
    planes[1] = psi.imag.reshape((2,) * n)
    new_psi, new_order = gate.apply(psi=planes, order=order)
    if any(x != y for x, y in zip(order, new_order)):
        raise RuntimeError("'order' has changed.")
    new_psi = np.asarray(new_psi)
    out = np.empty(2 ** n, dtype=state.complex_type)
    out.real = new_psi[0].reshape(-1)
    out.imag = new_psi[1].reshape(-1)
    state.upload(out)
