"""State-vector sharding across GPUs: one process per GPU, the top g = log2(p) index bits are
the rank, every rank holds 2^(n-g) amplitudes.

The reference has no counterpart: its evolution path refuses MPI
(/root/reference/hybridq/circuit/simulation/simulation.py:379-380) and
``examples/example-mpi.py:70-73`` replicates the evolution on every rank.  What is mirrored
is the bookkeeping idea of its host loop -- a logical->physical bit map that is only
restored at the end (simulation.py:512-513, :615-619, :655-663) -- lifted from "low bits of
one buffer" to "bits that are the rank".

Schedule (computed identically on every rank, pure Python, no communication):

  local     run every pending gate whose targets are all local bits (skipping over blocked
            gates when they commute, like the fuser) as fused tile passes;
  exchange  choose the new set of g rank bits = the logical bits whose next use is farthest
            away (Belady) among those that sit at a local position >= MIN_XPOS, and SWAP the s
            rank bits that must come in with the s local positions of the bits that go out --
            in place, wherever those positions are: no permutation pass moves them to the top
            first.  Every rank ships (2^s - 1)/2^s of its shard to the 2^s - 1 ranks that differ
            in the swapped rank bits.
  restore   at the end of the circuit put every logical bit back in place (<= 3 exchanges
            + one local permutation pass), so the result is in canonical order.

How an exchange moves data (round 2): FUSED into the write-back of the local pass that precedes
it.  Every rank maps its peers' second shard buffer over NVLink (cudaIpc through the C ABI) and the
last tile pass before the exchange streams each finished tile straight from shared memory to
where it belongs -- the own second buffer or a peer's (hq_plan_run_range_xchg, hq_ring_kernel<XCHG>)
-- so the NVLink traffic overlaps the gate arithmetic of the same pass tile by tile; one tiny
all-reduce afterwards is the cross-rank barrier, then the buffers swap roles.  (Round 1: a
permutation pass, then grouped NCCL send/recv of contiguous chunks, serialised with the kernels:
`exchange_sendrecv` below is still the path of engines without peer mapping -- the CPU tests.)

The local work is delegated to an *engine*; the product engine is :class:`CudaEngine`
(hand-written kernels through the C ABI).  The CPU-only tests inject an oracle-backed
engine to exercise this scheduling and the gloo send/recv pattern without a GPU -- the
product never does that.
"""
from __future__ import annotations

import math
import time
from typing import Sequence

import numpy as np

MIN_XPOS = 8          # lowest local bit position that may be swapped with a rank bit: the pieces that cross
                      # NVLink are contiguous runs of 2^MIN_XPOS amplitudes (2 KiB complex64)


# ------------------------------------------------------------------------------------------
# schedule
# ------------------------------------------------------------------------------------------
class Op:
    __slots__ = ("kind", "gates", "gate_ids", "perm", "gbits", "lpos")

    def __init__(self, kind, gates=None, gate_ids=None, perm=None, gbits=None, lpos=None):
        self.kind = kind            # 'local' | 'permute' | 'exchange'
        self.gates = gates          # [(U, physical positions)]
        self.gate_ids = gate_ids
        self.perm = perm            # local bit permutation: new bit i <- old bit perm[i]
        self.gbits = gbits          # exchange: rank-bit indices (0..g-1) ...
        self.lpos = lpos            # ... swapped one to one with these local bit positions


def plan_sharded(lowered: Sequence, n: int, g: int, restore: bool = True, min_xpos: int | None = None):
    """Turn [(U, logical positions)] into a list of Ops for p = 2^g ranks."""
    nl = n - g
    if nl < 1:
        raise ValueError(f"cannot shard a {n}-qubit state over 2^{g} ranks: no local qubits left")
    kmax = max((len(p) for _, p in lowered), default=0)
    if kmax > nl:
        # every target of a gate must be a local bit at the same time
        raise ValueError(f"a gate acts on {kmax} qubits but each of the 2^{g} ranks holds only {nl} local qubits; "
                         "use fewer ranks (or shard=False)")
    xmin = MIN_XPOS if min_xpos is None else min_xpos
    xmin = max(0, min(xmin, nl - g - kmax))       # small shards: fall back to lower positions
    where = list(range(n))                      # logical bit -> physical bit (>= nl: rank bit)
    ops: list[Op] = []
    pending = list(range(len(lowered)))
    stats = {"exchanges": 0, "exchange_bits": 0, "permutes": 0, "moved_shard_fraction": 0.0,
             "crossing_gates": sum(1 for _, p in lowered if any(b >= nl for b in p)),
             "gates": len(lowered)}

    while pending:
        blocked: set[int] = set()
        run, rest = [], []
        for gi in pending:
            bits = lowered[gi][1]
            if any(b in blocked for b in bits) or any(where[b] >= nl for b in bits):
                blocked.update(bits)
                rest.append(gi)
            else:
                run.append(gi)
        if run:
            ops.append(Op("local", gates=[(lowered[gi][0], [where[b] for b in lowered[gi][1]]) for gi in run],
                          gate_ids=run))
        pending = rest
        if not pending:
            break
        first_use = {}
        for idx, gi in enumerate(pending):
            for b in lowered[gi][1]:
                first_use.setdefault(b, idx)
        need_now = set(lowered[pending[0]][1])
        # candidates for being a rank bit: the current ones, and local bits high enough to be swapped out;
        # farthest next use first; ties: the bits whose home is a rank position (so that the final restore
        # finds them in place), then the current rank bits, then higher physical positions
        cands = [b for b in range(n) if b not in need_now and (where[b] >= nl or where[b] >= xmin)]
        order = sorted(cands, key=lambda b: (-first_use.get(b, math.inf), b < nl, where[b] < nl, -where[b]))
        before = list(where)
        emit_exchange(ops, stats, where, n, nl, order[:g])
        if where == before:
            raise RuntimeError("sharded schedule made no progress (internal error): "
                               f"pending gate on bits {lowered[pending[0]][1]}, rank bits "
                               f"{[b for b in range(n) if where[b] >= nl]}")

    if restore and g > 0:
        for _ in range(3):
            inv = {where[b]: b for b in range(n)}
            wrong_pos = [pos for pos in range(nl, n) if inv[pos] != pos]      # rank positions with a wrong occupant
            if not wrong_pos:
                break
            # a wrong occupant is swapped with its position's rightful bit when that one is local now, otherwise
            # with any free high local position (the next round then finds the rightful bit local)
            reserved = {where[pos] for pos in wrong_pos if where[pos] < nl}
            taken: set[int] = set()
            gb, lp = [], []
            for pos in wrong_pos:
                src = where[pos]
                if src >= nl or src in taken:
                    src = next(p for p in range(nl - 1, -1, -1) if p not in taken and p not in reserved)
                taken.add(src)
                gb.append(pos - nl)
                lp.append(src)
            _apply_swap(ops, stats, where, n, nl, gb, lp)
        if any(where[i] != i for i in range(nl)):
            # new bit i <- old bit where[i] (logical bit i currently lives at physical where[i])
            ops.append(Op("permute", perm=[where[i] for i in range(nl)]))
            stats["permutes"] += 1
            for i in range(nl):
                where[i] = i
    return ops, stats, where


def _apply_swap(ops, stats, where, n, nl, gbits, lpos):
    """Record the exchange that swaps rank position nl + gbits[j] with local position lpos[j] and update
    `where` (logical bit -> physical position)."""
    s = len(gbits)
    if s == 0:
        return
    inv = [0] * n
    for b in range(n):
        inv[where[b]] = b
    # the kernels take ascending rank-bit order (digit bit j <-> gbits[j])
    pairs = sorted(zip(gbits, lpos))
    gbits = [x for x, _ in pairs]
    lpos = [y for _, y in pairs]
    ops.append(Op("exchange", gbits=gbits, lpos=lpos))
    stats["exchanges"] += 1
    stats["exchange_bits"] += s
    stats["moved_shard_fraction"] += (2 ** s - 1) / 2 ** s
    for gb, lp in zip(gbits, lpos):
        b_rank, b_loc = inv[nl + gb], inv[lp]
        where[b_rank], where[b_loc] = lp, nl + gb


def emit_exchange(ops, stats, where, n, nl, new_global):
    """Exchange so that the rank bits become the logical bits `new_global`: every rank position whose
    occupant is not in `new_global` is swapped with the local position of an incoming bit, placing logical bit
    b at rank position b whenever b >= nl is incoming and that position is being vacated."""
    cur_global = [b for b in range(n) if where[b] >= nl]
    s_out = sorted([b for b in cur_global if b not in new_global], key=lambda b: where[b])
    s_in = [b for b in new_global if b not in cur_global]
    s = len(s_out)
    if s == 0:
        return
    slots = [where[b] for b in s_out]                 # physical rank positions being vacated
    assigned = [None] * s
    rest = []
    for b in s_in:
        if b >= nl and b in slots:
            assigned[slots.index(b)] = b              # home position
        else:
            rest.append(b)
    for j in range(s):
        if assigned[j] is None:
            assigned[j] = rest.pop(0)
    _apply_swap(ops, stats, where, n, nl, [p - nl for p in slots], [where[b] for b in assigned])


# ------------------------------------------------------------------------------------------
# exchange without peer mapping (gloo CPU tests, GPUs without P2P): send/recv of gathered pieces
# ------------------------------------------------------------------------------------------
def exchange_sendrecv(dist, src, dst, rank: int, gbits: Sequence[int], lpos: Sequence[int]):
    """Swap rank bits `gbits` with the local index bits `lpos` (one to one).  `src`, `dst`: 1-D tensors holding
    the shard; the result is written to `dst`.  The amplitudes whose local bits `lpos` spell the digit D go to
    the rank whose bits `gbits` equal D and land there at the same local index with those bits replaced by the
    sender's own digit.  Pieces are gathered into contiguous buffers for dist.isend / irecv."""
    import torch
    s = len(gbits)
    mine = 0
    for j, gb in enumerate(gbits):
        mine |= ((rank >> gb) & 1) << j
    idx = torch.arange(src.numel(), device=src.device)
    digit = torch.zeros_like(idx)
    for j, lp in enumerate(lpos):
        digit |= ((idx >> lp) & 1) << j
    sel = [torch.nonzero(digit == D).reshape(-1) for D in range(1 << s)]
    ops, recv = [], {}
    for D in range(1 << s):
        if D == mine:
            continue
        partner = rank
        for j, gb in enumerate(gbits):
            partner = (partner & ~(1 << gb)) | (((D >> j) & 1) << gb)
        out = src[sel[D]].contiguous()
        recv[D] = torch.empty_like(out)
        ops.append(dist.P2POp(dist.isend, out, partner))
        ops.append(dist.P2POp(dist.irecv, recv[D], partner))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    dst[sel[mine]] = src[sel[mine]]
    for r in reqs:
        r.wait()
    # what partner(D) sent are ITS amplitudes with local digit == mine; they land where my local digit is D
    for D, buf in recv.items():
        dst[sel[D]] = buf


# ------------------------------------------------------------------------------------------
# engines
# ------------------------------------------------------------------------------------------
class CudaEngine:
    """Local work on the GPU through libhybridq_b200.so."""

    def __init__(self, n_local: int, ctype, plan_options=None):
        import hybridq_b200 as hb
        self.hb = hb
        self.n_local = n_local
        self.ctype = np.dtype(ctype)
        self.plan_options = plan_options
        self._plans = {}

    def alloc(self):
        # own cudaMalloc blocks: peers map them with cudaIpc for the fused exchange
        return self.hb.DeviceState(self.n_local, self.ctype, ipc=True)

    def tensor(self, st):
        return st.tensor

    def _plan(self, key, gates):
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self.hb.Plan(gates, self.n_local, self.ctype, self.plan_options)
        return plan

    def run_gates(self, st, key, gates):
        plan = self._plan(key, gates)
        plan.run(st)
        return plan.n_passes

    # -- fused exchange (peer buffers mapped over NVLink) -------------------------------------------------
    def map_peers(self, dist, buffers):
        """Exchange cudaIpc handles of this rank's shard buffers with every rank and map the peers' buffers.
        Returns ptrs[buffer index][rank] (own buffers: their own pointers) or None when peer mapping is not
        possible (then the runner falls back to send/recv)."""
        from .state import open_ipc
        rank, world = dist.get_rank(), dist.get_world_size()
        try:
            mine = [b.raw.ipc_handle() for b in buffers]
        except Exception:
            mine = None
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if any(h is None for h in gathered):
            return None
        ptrs = [[0] * world for _ in buffers]
        ok = 1
        try:
            for r in range(world):
                for i, b in enumerate(buffers):
                    ptrs[i][r] = b.ptr.value if r == rank else open_ipc(gathered[r][i], b.device)
        except Exception:
            ok = 0
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        return ptrs if all(flags) else None

    def run_gates_xchg(self, st, key, gates, digit, lpos, dst_ptrs):
        """Local gates (may be empty) with the exchange fused into the write-back of the last pass."""
        plan = self._plan(key, gates)
        plan.run_xchg(st, digit, lpos, dst_ptrs)
        return max(1, plan.n_passes)

    def permute(self, st, key, perm):
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self.hb.BitPermPlan(perm, self.n_local, self.ctype)
        plan.run(st)
        return plan.n_passes

    def init_random(self, st, seed, index_offset):
        st.init_random(seed=seed, index_offset=index_offset, scale=1.0)

    def init_product(self, st, spec_local, factor):
        st.init_product(spec_local)
        if factor != 1.0:
            st.scale(factor)

    def upload(self, st, host):
        st.upload(host)

    def download(self, st, out=None):
        return st.download(out)

    def norm2(self, st):
        return st.norm2()

    def scale(self, st, f):
        st.scale(f)

    def marginal(self, st, pos):
        return st.marginal(pos)

    def project(self, st, pos, outcome, scale_re, scale_im):
        st.project(pos, outcome, scale_re, scale_im)

    def sync(self):
        import torch
        torch.cuda.synchronize()


class ShardedRunner:
    """Executes a sharded schedule; one instance per rank."""

    def __init__(self, n: int, lowered: Sequence, ctype, dist, engine=None, plan_options=None, restore=True,
                 key=None):
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.g = int(round(math.log2(self.world)))
        if 2 ** self.g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n = n
        self.n_local = n - self.g
        self.ctype = np.dtype(ctype)
        self._schedules = {}                # circuit key -> (ops, stats, final_where)
        self._segment = key if key is not None else 0
        self.ops, self.stats, self.final_where = plan_sharded(lowered, n, self.g, restore=restore)
        if key is not None:
            self._schedules[key] = (self.ops, dict(self.stats), self.final_where)
        self.n_gates = len(lowered)
        self.engine = engine if engine is not None else CudaEngine(self.n_local, ctype, plan_options)
        self.a = self.engine.alloc()
        self.b = self.engine.alloc()
        # fused exchange: map every rank's two shard buffers; peer[i][r] = pointer to rank r's buffer i
        self._bufs = [self.a, self.b]
        self._cur = 0                       # which buffer holds the state (the same on every rank)
        self.peer = None
        if self.g > 0 and hasattr(self.engine, "map_peers"):
            self.peer = self.engine.map_peers(dist, self._bufs)
        self.fused = self.peer is not None
        self._flag = None
        self.local_passes = 0
        self.exchange_ms = 0.0
        self._count_passes = True
        self.complex_type = self.ctype

    def replan(self, lowered: Sequence, restore: bool = True, key=None, accumulate: bool = True):
        """Swap in the schedule of another gate list (the next segment of a circuit that is interrupted by
        measurement gates, or another circuit on a reused runner); the shard buffers and their contents stay.
        `key` (hashable) identifies the gate list: schedules and the engine's compiled plans are cached under
        it, so running the same circuit again costs no planning."""
        if key is not None and key in self._schedules:
            self.ops, stats, self.final_where = self._schedules[key]
        else:
            self.ops, stats, self.final_where = plan_sharded(lowered, self.n, self.g, restore=restore)
            if key is not None:
                self._schedules[key] = (self.ops, dict(stats), self.final_where)
        if accumulate:
            for k, v in stats.items():
                self.stats[k] = self.stats.get(k, 0) + v
            self.n_gates += len(lowered)
        else:
            self.stats = dict(stats)
            self.n_gates = len(lowered)
        self._segment = key if key is not None else (self._segment + 1 if isinstance(self._segment, int) else 1)

    def close(self):
        """Unmap the peers' buffers and release this rank's (collective: every rank must call it)."""
        from ._lib import lib
        import ctypes
        if self.peer is not None:
            for table in self.peer:
                for r, p in enumerate(table):
                    if r != self.rank and p:
                        lib.hq_ipc_close(ctypes.c_void_p(p))
            self.peer = None
            self.fused = False
        self.engine.sync()
        self.dist.barrier()
        self.a = self.b = None
        self._bufs = []
        self.engine._plans.clear() if hasattr(self.engine, "_plans") else None

    # -- measurement support on the sharded state (canonical bit order required: call between segments) ------
    def marginal(self, pos: Sequence[int]):
        """(2^k, 2) sums of re^2 / im^2 per outcome of the LOGICAL index bits `pos` (bit j of the outcome =
        index bit pos[j]), over the whole state: local reduction per rank, the rank bits contribute this rank's
        own bit values, one all-reduce."""
        import torch
        if any(w != i for i, w in enumerate(self.final_where)):
            raise RuntimeError("marginal() needs the canonical bit order (plan with restore=True)")
        nl = self.n_local
        local = [(j, p) for j, p in enumerate(pos) if p < nl]
        loc = np.asarray(self.engine.marginal(self.a, [p for _, p in local]), dtype=np.float64)
        fixed = sum((((self.rank >> (p - nl)) & 1) << j) for j, p in enumerate(pos) if p >= nl)
        full = np.zeros((2 ** len(pos), 2), dtype=np.float64)
        for s_loc in range(2 ** len(local)):
            s = fixed
            for i, (j, _) in enumerate(local):
                s |= ((s_loc >> i) & 1) << j
            full[s] += loc[s_loc]
        t = torch.from_numpy(full).to(self.engine.tensor(self.a).device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def project(self, pos: Sequence[int], outcome: int, scale_re: float = 1.0, scale_im: float = 1.0):
        """Keep the amplitudes whose LOGICAL index bits `pos` spell `outcome` (scaled plane-wise), zero the rest;
        a rank whose own bits contradict the outcome zeroes its whole shard."""
        nl = self.n_local
        for j, p in enumerate(pos):
            if p >= nl and ((self.rank >> (p - nl)) & 1) != ((outcome >> j) & 1):
                self.engine.project(self.a, [], 0, 0.0, 0.0)
                return self
        local = [(j, p) for j, p in enumerate(pos) if p < nl]
        out_loc = sum((((outcome >> j) & 1) << i) for i, (j, _) in enumerate(local))
        self.engine.project(self.a, [p for _, p in local], out_loc, scale_re, scale_im)
        return self

    def broadcast_int(self, value: int, src: int = 0) -> int:
        import torch
        t = torch.tensor([int(value)], dtype=torch.int64, device=self.engine.tensor(self.a).device)
        self.dist.broadcast(t, src)
        return int(t.item())

    def describe(self):
        s = self.stats
        return (f"{self.world} GPUs, top {self.g} index bits = rank; {s['crossing_gates']}/{s['gates']} gates touch a "
                f"sharded qubit; {s['exchanges']} exchanges ({s['exchange_bits']} bit swaps, "
                f"{s['moved_shard_fraction']:.2f} shards moved per GPU, "
                f"{'fused into the preceding tile pass, peer stores over NVLink' if self.fused else 'send/recv'}), "
                f"{s['permutes']} local permutation passes, {self.local_passes} tile passes per step")

    def init_state(self, seed: int):
        self.engine.init_random(self.a, seed, self.rank * (2 ** self.n_local))
        n2 = self.engine.norm2(self.a)
        tot = self._allreduce_sum(n2)
        self.engine.scale(self.a, 1.0 / math.sqrt(tot))

    def init_product(self, spec: str):
        """Product state from a string of n characters 0,1,+,- (spec[0] = most significant bit):
        the rank bits select a scalar factor, the local bits a local product state."""
        if len(spec) == 1:
            spec = spec * self.n
        if len(spec) != self.n:
            raise ValueError("Wrong number of qubits for initial/final state.")
        r = 2 ** -0.5
        factor = 1.0
        for j in range(self.g):                      # spec[j] <-> index bit n-1-j = rank bit g-1-j
            bit = (self.rank >> (self.g - 1 - j)) & 1
            factor *= {"0": (1, 0), "1": (0, 1), "+": (r, r), "-": (r, -r)}[spec[j]][bit]
        self.engine.init_product(self.a, spec[self.g:], factor)

    def dump(self, path, **kw):
        """Sharded checkpoint: every rank writes `path.rank<r>of<w>` (+ .json) with its shard in canonical order
        (DeviceState.dump: chunked through pinned memory, so 64 GiB shards do not need 64 GiB of host RAM)."""
        if any(w != i for i, w in enumerate(self.final_where)):
            raise RuntimeError("dump() needs the canonical bit order (plan with restore=True)")
        self.a.dump(f"{path}.rank{self.rank}of{self.world}",
                    meta={"rank": self.rank, "world": self.world, "n_qubits_global": self.n}, **kw)
        self.dist.barrier()

    def load(self, path, **kw):
        """Read a checkpoint written by :meth:`dump` with the same number of ranks."""
        self.a.load(f"{path}.rank{self.rank}of{self.world}",
                    expect={"rank": self.rank, "world": self.world, "n_qubits_global": self.n}, **kw)
        self.dist.barrier()

    def load_shard(self, host_shard):
        self.engine.upload(self.a, host_shard)

    def download_shard(self, out=None):
        return self.engine.download(self.a, out)

    def _allreduce_sum(self, v: float) -> float:
        import torch
        t = self.engine.tensor(self.a)
        x = torch.tensor([v], dtype=torch.float64, device=t.device)
        self.dist.all_reduce(x)
        return float(x.item())

    def norm2(self) -> float:
        return self._allreduce_sum(self.engine.norm2(self.a))

    def _digit_and_dst(self, op):
        """This rank's digit over the swapped rank bits and, for every digit D, the destination buffer: the
        not-current buffer of the rank whose bits op.gbits spell D."""
        mine = 0
        for j, gb in enumerate(op.gbits):
            mine |= ((self.rank >> gb) & 1) << j
        dst = []
        for D in range(1 << len(op.gbits)):
            partner = self.rank
            for j, gb in enumerate(op.gbits):
                partner = (partner & ~(1 << gb)) | (((D >> j) & 1) << gb)
            dst.append(self.peer[1 - self._cur][partner])
        return mine, dst

    def _barrier_and_swap(self):
        """After a fused exchange: every rank must have finished writing into everybody's buffers before anyone
        reads its own -- one 1-element all-reduce in stream order -- then the two buffers swap roles."""
        import torch
        if self._flag is None:
            self._flag = torch.zeros(1, dtype=torch.int32, device=self.engine.tensor(self.a).device)
        self.dist.all_reduce(self._flag)
        self.a, self.b = self.b, self.a
        self._cur ^= 1

    def step(self, time_exchange: bool = False):
        passes = 0
        ops = self.ops
        i = 0
        while i < len(ops):
            op = ops[i]
            nxt = ops[i + 1] if i + 1 < len(ops) else None
            if self.fused and op.kind == "local" and nxt is not None and nxt.kind == "exchange":
                mine, dst = self._digit_and_dst(nxt)
                passes += self.engine.run_gates_xchg(self.a, ("g", self._segment, i), op.gates, mine, nxt.lpos, dst)
                self._barrier_and_swap()
                i += 2
                continue
            if op.kind == "local":
                passes += self.engine.run_gates(self.a, ("g", self._segment, i), op.gates)
            elif op.kind == "permute":
                passes += self.engine.permute(self.a, ("p", self._segment, i), op.perm)
            elif self.fused:
                mine, dst = self._digit_and_dst(op)
                passes += self.engine.run_gates_xchg(self.a, ("x", self._segment, i), [], mine, op.lpos, dst)
                self._barrier_and_swap()
            else:
                if time_exchange:
                    self.engine.sync()
                    t0 = time.perf_counter()
                exchange_sendrecv(self.dist, self.engine.tensor(self.a), self.engine.tensor(self.b), self.rank,
                                  op.gbits, op.lpos)
                self.a, self.b = self.b, self.a
                self._cur ^= 1
                if time_exchange:
                    self.engine.sync()
                    self.exchange_ms += 1e3 * (time.perf_counter() - t0)
            i += 1
        self.local_passes = passes

    def kernel_time_ms(self, reps: int = 2) -> float:
        """Device time of the local launches only (no exchanges), CUDA events."""
        import torch
        total = 0.0
        for _ in range(reps):
            for i, op in enumerate(self.ops):
                if op.kind == "exchange":
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if op.kind == "local":
                    self.engine.run_gates(self.a, ("g", self._segment, i), op.gates)
                else:
                    self.engine.permute(self.a, ("p", self._segment, i), op.perm)
                e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
        return total / reps

    def extra_roofline(self, peak):
        return {}

    def gather(self):
        """Full state on every rank (tests / small n only), canonical order required."""
        import torch
        t = self.engine.tensor(self.a)
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return torch.cat(out).cpu().numpy()

    def e2e(self, gates, steps, barrier, dist):
        """Pinned host shard in -> circuit -> pinned host shard out, per rank."""
        import torch
        t = self.engine.tensor(self.a)
        host_in = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        host_out = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        host_in.zero_()
        if self.rank == 0:
            host_in[0] = 1
        for it in range(steps + 1):
            if it == 1:
                barrier()
                t0 = time.perf_counter()
            self.engine.tensor(self.a).copy_(host_in, non_blocking=True)
            self.step()
            host_out.copy_(self.engine.tensor(self.a), non_blocking=True)
            torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        x = torch.tensor([dt], dtype=torch.float64, device=t.device)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        dt = float(x.item())
        nbytes = t.numel() * t.element_size()
        return {"value": self.n_gates * steps / dt, "unit": "gate-applies/s",
                "h2d_bytes_per_step": nbytes * self.world, "d2h_bytes_per_step": nbytes * self.world,
                "ms_per_step": 1e3 * dt / steps, "steps": steps,
                "api": "hybridq_b200.dist.ShardedRunner: pinned host shard -> HBM, step(), HBM -> pinned host shard"}
