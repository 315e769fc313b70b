"""Seeded synthetic circuits for parity tests and benchmarks.

A circuit here is a plain list of :class:`GateApply` records -- one dense
2^k x 2^k complex matrix on k qubit labels -- which is exactly what the reference's
hot loop consumes from a ``Gate`` (``gate.qubits`` and ``gate.matrix()``,
/root/reference/hybridq/circuit/simulation/simulation.py:633-637).  ``GateApply``
duck-types that part of the reference Gate API, so the same objects can be handed
to :func:`hybridq_b200.simulate` or (wrapped by ``to_hybridq``) to the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Sequence

import numpy as np


@dataclass
class GateApply:
    """Minimal stand-in for ``hybridq.gate.MatrixGate``: ``qubits[0]`` indexes the
    most significant bit of the matrix row/column index (reference convention,
    simulation.py:633 reverses ``gate.qubits`` to get LSB-first positions)."""
    U: np.ndarray
    qubits: tuple
    name: str = "MATRIX"
    tags: dict = field(default_factory=dict)

    def matrix(self, order=None) -> np.ndarray:
        if order is not None and tuple(order) != tuple(self.qubits):
            raise NotImplementedError("GateApply.matrix(order=...) is not supported")
        return self.U

    @property
    def n_qubits(self) -> int:
        return len(self.qubits)

    def provides(self, what) -> bool:
        what = [what] if isinstance(what, str) else list(what)
        return all(w in ("qubits", "matrix", "n_qubits", "name", "tags") for w in what)


@dataclass
class ProjectionApply:
    """Stand-in for ``hybridq.gate.Projection`` (gate/projection.py:122): project the qubits onto the z-basis
    state ``state`` (a string of 0/1, one character per qubit) and renormalise.  Carries no ``apply``: both this
    and the reference's ProjectionGate run on the device inside :func:`hybridq_b200.simulate`."""
    qubits: tuple
    state: str
    name: str = "PROJECTION"
    tags: dict = field(default_factory=dict)


@dataclass
class MeasureApply:
    """Stand-in for ``hybridq.gate.Measure`` (gate/measure.py:122): sample an outcome of the qubits with
    ``numpy.random.choice`` (the reference's generator, so seeding numpy reproduces its draw), project onto it
    and renormalise."""
    qubits: tuple
    name: str = "MEASURE"
    tags: dict = field(default_factory=dict)


def haar_unitary(dim: int, rng: np.random.Generator) -> np.ndarray:
    from scipy.stats import unitary_group
    if dim == 1:
        return np.exp(2j * np.pi * rng.random((1, 1)))
    return unitary_group.rvs(dim, random_state=rng)


def matching_circuit(n: int, depth: int = 20, seed: int | None = None,
                     p_two: float = 0.5) -> list[GateApply]:
    """SURVEY.md §8d config 1/2 generator: `depth` layers; each layer is a random
    perfect matching of the n qubits; each pair gets, with probability `p_two`, one
    Haar 4x4 gate, otherwise two Haar 2x2 gates.  Qubit labels are ints 0..n-1 and
    label q maps to index bit n-1-q (first sorted qubit = MSB, simulation.py:512)."""
    rng = np.random.default_rng(n if seed is None else seed)
    gates: list[GateApply] = []
    for _ in range(depth):
        perm = rng.permutation(n)
        for a, b in zip(perm[0::2], perm[1::2]):
            if rng.random() < p_two:
                gates.append(GateApply(haar_unitary(4, rng), (int(a), int(b))))
            else:
                gates.append(GateApply(haar_unitary(2, rng), (int(a),)))
                gates.append(GateApply(haar_unitary(2, rng), (int(b),)))
        if n % 2:
            gates.append(GateApply(haar_unitary(2, rng), (int(perm[-1]),)))
    return gates


def ksweep_circuit(n: int, k: int, n_gates: int = 20, seed: int | None = None,
                   bits: Sequence[int] | None = None) -> list[GateApply]:
    """`n_gates` Haar 2^k x 2^k gates on k distinct
    uniformly-random qubits (or drawn from `bits`)."""
    rng = np.random.default_rng(33 * 10 + k if seed is None else seed)
    pool = np.arange(n) if bits is None else np.asarray(bits)
    return [GateApply(haar_unitary(2 ** k, rng),
                      tuple(int(q) for q in rng.permutation(pool)[:k]))
            for _ in range(n_gates)]


def sharded_circuit(n: int, n_global: int, depth: int = 20, frac_global: float = 0.2,
                    seed: int | None = None) -> list[GateApply]:
    """matching circuit whose pairs are re-drawn so that about
    `frac_global` of the gate-applies touch one of the `n_global` most significant
    qubits (labels 0..n_global-1, the sharded ones)."""
    rng = np.random.default_rng(n if seed is None else seed)
    gates: list[GateApply] = []
    glob = list(range(n_global))
    loc = list(range(n_global, n))
    per_layer = n // 2
    for _ in range(depth):
        for _ in range(per_layer):
            if rng.random() < frac_global and glob:
                a = int(rng.choice(glob))
                b = int(rng.choice(loc))
            else:
                a, b = (int(x) for x in rng.choice(loc, size=2, replace=False))
            if rng.random() < 0.5:
                gates.append(GateApply(haar_unitary(4, rng), (a, b)))
            else:
                gates.append(GateApply(haar_unitary(2, rng), (a,)))
                gates.append(GateApply(haar_unitary(2, rng), (b,)))
    return gates


def to_positions(gates: Sequence, qubits: Sequence | None = None):
    """Lower Gate-like objects to ``[(U, pos)]`` with LSB-first index-bit positions,
    exactly as the reference does: ``_map[q] = n-1-index(q)`` over the sorted qubit
    list and ``pos = [_map[q] for q in reversed(gate.qubits)]``
    (simulation.py:512-513, :633)."""
    if qubits is None:
        qubits = sorted({q for g in gates for q in g.qubits})
    n = len(qubits)
    qmap = {q: n - 1 - i for i, q in enumerate(qubits)}
    return [(np.asarray(g.matrix()), [qmap[q] for q in reversed(tuple(g.qubits))])
            for g in gates], n


def random_state(n: int, dtype="complex64", seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    ft = np.float32 if np.dtype(dtype) == np.complex64 else np.float64
    psi = rng.standard_normal(2 ** n, dtype=ft) + 1j * rng.standard_normal(2 ** n, dtype=ft)
    psi /= np.linalg.norm(psi.astype(np.complex128))
    return psi.astype(dtype)
