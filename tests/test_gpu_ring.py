"""The pipelined ring kernel (hq_ring_kernel: producer warp + two consumer groups over a 3-stage shared-memory
ring, hybridq_b200/csrc/hq_kernels.cu) against the oracle.  On large states it is the default; here it is forced
onto small ones with hq_set_ring(1) so that the oracle finishes in seconds and the edge cases are hit: fewer tiles
than stages, fewer tiles than SMs, every kernel class, the permuted drain, and the exchange redirect of the
write-back that the multi-GPU path uses (checked on ONE GPU by playing all ranks in turn)."""
import ctypes

import numpy as np
import pytest

from helpers import TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb():
    import hybridq_b200
    return hybridq_b200


@pytest.fixture()
def ring(hb):
    hb.lib.hq_set_ring(1)
    yield
    hb.lib.hq_set_ring(-1)


def _rand_state(rng, n, ctype):
    psi = (rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)).astype(ctype)
    return (psi / np.linalg.norm(psi)).astype(ctype)


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
@pytest.mark.parametrize("n", [13, 14, 16, 19, 22])
def test_ring_kernel_circuit_vs_oracle(hb, ring, oracle, c_oracle, ctype, n):
    """n = 13 / 14: one or two tiles (fewer than the ring has stages); 16: 8 tiles; 19 / 22: fewer / more tiles
    than SMs x stages.  Gates of every k = 1..6 so that all kernel classes run."""
    from hybridq_b200.circuits import matching_circuit, to_positions, haar_unitary
    rng = np.random.default_rng(100 + n)
    lowered, _ = to_positions(matching_circuit(n, depth=6, seed=n), qubits=list(range(n)))
    for k in (3, 4, 5, 6):
        lowered.insert(int(rng.integers(0, len(lowered))), (haar_unitary(2 ** k, rng), [int(x) for x in rng.permutation(n)[:k]]))
    psi = _rand_state(rng, n, ctype)
    ref = oracle.evolve_oracle(psi, [(U.astype(ctype), p) for U, p in lowered], c_oracle)
    for opts in (None, hb.PlanOptions(mma_min_k=0), hb.PlanOptions(mma_min_k=2), hb.PlanOptions(fuse=0)):
        st = hb.DeviceState(n, ctype).upload(psi)
        launches0 = hb.lib.hq_launch_count()
        hb.Plan(lowered, n, ctype, opts).run(st)
        assert hb.lib.hq_launch_count() > launches0
        assert np.abs(st.download() - ref).max() <= TOL[ctype], (n, ctype)


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
def test_ring_kernel_bit_permutation_is_bit_exact(hb, ring, oracle, ctype):
    rng = np.random.default_rng(5)
    n = 17
    psi = _rand_state(rng, n, ctype)
    for _ in range(4):
        perm = rng.permutation(n)
        out = hb.DeviceState(n, ctype).upload(psi).permute_bits(perm).download()
        assert np.array_equal(out, oracle.numpy_swap(psi, perm))


@pytest.mark.parametrize("ctype", ["complex64", "complex128"])
@pytest.mark.parametrize("s,pos", [(1, [15]), (2, [9, 14]), (3, [7, 11, 15]), (2, [13, 3])])
def test_exchange_redirect_single_gpu_model(hb, oracle, c_oracle, ctype, s, pos):
    """hq_plan_run_range_xchg: 2^s 'ranks' played one after the other on one GPU.  Each runs the same local gates
    on its shard and writes the result through the redirect into the ranks' second buffers; afterwards the global
    state must be: local gates applied, then rank bit j swapped with local bit pos[j]."""
    from hybridq_b200.circuits import matching_circuit, to_positions
    rng = np.random.default_rng(31 + s)
    nl = 16
    ranks = 1 << s
    n = nl + s
    psi = _rand_state(rng, n, ctype)
    lowered, _ = to_positions(matching_circuit(nl, depth=3, seed=s), qubits=list(range(nl)))
    a = [hb.DeviceState(nl, ctype).upload(psi[r << nl:(r + 1) << nl]) for r in range(ranks)]
    b = [hb.DeviceState(nl, ctype) for _ in range(ranks)]
    for t in b:
        t.tensor.zero_()
    plan = hb.Plan(lowered, nl, ctype)
    dst = (ctypes.c_void_p * 8)(*([t.ptr.value for t in b] + [None] * (8 - ranks)))
    posc = (ctypes.c_uint32 * 4)(*(pos + [0] * (4 - s)))
    import torch
    for r in range(ranks):
        hb._lib.check(hb.lib.hq_plan_run_range_xchg(plan._h, a[r].ptr, 0, plan.n_passes, s, r, posc, dst, None), "xchg")
    torch.cuda.synchronize()
    # model: local gates on the global vector, then swap global bit nl + j with local bit pos[j]
    ref = oracle.evolve_oracle(psi, [(U.astype(ctype), p) for U, p in lowered], c_oracle)
    perm = list(range(n))
    for j, p in enumerate(pos):
        perm[p], perm[nl + j] = nl + j, p
    want = oracle.numpy_swap(ref, perm)
    got = np.concatenate([t.download() for t in b])
    assert np.abs(got - want).max() <= TOL[ctype]
    # a gate-less redirect (pure exchange) is bit-exact
    empty = hb.Plan([], nl, ctype)
    for t in b:
        t.tensor.zero_()
    for r in range(ranks):
        a[r].upload(psi[r << nl:(r + 1) << nl])
        hb._lib.check(hb.lib.hq_plan_run_range_xchg(empty._h, a[r].ptr, 0, empty.n_passes, s, r, posc, dst, None), "xchg")
    torch.cuda.synchronize()
    got = np.concatenate([t.download() for t in b])
    assert np.array_equal(got, oracle.numpy_swap(psi, perm))
