#!/usr/bin/env python
"""Small launch sequence for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family on small
states -- tile kernel (FFMA2 slots k = 1..3 with and without bit 0, generic paths, tensor cores, scalar + rank one,
two-phase k = 7), ring kernel, exchange redirect, direct kernel, permutation, marginal / project, init / reductions.
Results are checked against the oracle so that a sanitizer-clean run is also a correct one."""
import ctypes
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import hybridq_b200 as hb  # noqa: E402
from hybridq_b200.circuits import matching_circuit, to_positions, haar_unitary, random_state  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(0)
n = 15
for ctype, tol in (("complex64", 1e-6), ("complex128", 1e-12)):
    lowered, _ = to_positions(matching_circuit(n, depth=3, seed=5), qubits=list(range(n)))
    for k in (1, 3, 4, 5, 7):
        lowered.append((haar_unitary(2 ** k, rng), [int(x) for x in rng.permutation(n)[:k]]))
    lowered.append((haar_unitary(8, rng), [0, 4, 9]))
    u = rng.standard_normal(16) + 1j * rng.standard_normal(16)
    lowered.append((0.95 * np.eye(16) + 0.01 * np.outer(u, u.conj()), [2, 5, 8, 11]))
    psi = random_state(n, ctype, seed=1)
    ref = O.evolve_oracle(psi, [(U.astype(ctype), p) for U, p in lowered])
    for ring in (0, 1):
        hb.lib.hq_set_ring(ring)
        for opts in (None, hb.PlanOptions(mma_min_k=0), hb.PlanOptions(fuse=0), hb.PlanOptions(tile_bits=11, mma_min_k=2)):
            st = hb.DeviceState(n, ctype).upload(psi)
            hb.Plan(lowered, n, ctype, opts).run(st)
            err = float(np.abs(st.download() - ref).max())
            assert err <= tol, (ctype, ring, err)
    hb.lib.hq_set_ring(-1)
    # exchange redirect, two 'ranks' on one GPU
    nl = n - 1
    a = [hb.DeviceState(nl, ctype).upload(psi[r << nl:(r + 1) << nl]) for r in range(2)]
    b = [hb.DeviceState(nl, ctype) for _ in range(2)]
    low2, _ = to_positions(matching_circuit(nl, depth=2, seed=2), qubits=list(range(nl)))
    plan = hb.Plan(low2, nl, ctype)
    for r in range(2):
        plan.run_xchg(a[r], r, [10], [b[0].ptr.value, b[1].ptr.value])
    torch.cuda.synchronize()
    want = O.evolve_oracle(psi, [(U.astype(ctype), p) for U, p in low2])
    perm = list(range(n))
    perm[10], perm[nl] = nl, 10
    want = O.numpy_swap(want, perm)
    got = np.concatenate([t.download() for t in b])
    assert np.abs(got - want).max() <= tol
    # permutation, measurement support, reductions
    st = hb.DeviceState(n, ctype).upload(psi)
    st.permute_bits(rng.permutation(n))
    st.marginal([0, 3, 7])
    st.marginal(list(range(12)))
    st.project([1, 4], 2, 1.0, 1.0)
    st.norm2()
    hb.DeviceState(n, ctype).init_product("+-01" * 3 + "0+1").norm2()
    hb.DeviceState(n, ctype).init_random(seed=3).sample(100, seed=1)
print("SANITIZE_TARGET_OK")
