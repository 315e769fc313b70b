"""Subprocess worker of tests/test_gpu_reference_dropin.py: runs the UNMODIFIED reference package
(oracle/_ref/pkg/hybridq, a git-ignored copy of /root/reference/hybridq made by oracle/Makefile) on SURVEY.md
8(d) config 1 -- 20-qubit depth-20 random matching circuit, optimize='evolution' -- and stores the final
states.  Which native core the reference binds is decided by the caller through LD_LIBRARY_PATH
(hybridq/utils/utils.py:535-553 resolves 'hybridq.so' by bare name): this repo's drop-in libraries or the
reference's own AVX2 build.

    python ref_simulate_worker.py <out.npz> <mode>      mode = plain | dispatch

`dispatch` additionally applies INTEGRATION.md section 2 as a monkey patch (the reference files stay untouched):
the hybridq branch of `_simulate_evolution` hands the pre-processed circuit to hybridq_b200.simulate, so the
state stays in HBM for the whole gate loop.

The circuit generator is loaded from hybridq_b200/circuits.py BY FILE PATH: importing the hybridq_b200 package
would dlopen libhybridq_b200.so, which must not happen in the arm that runs on the reference core.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def load_circuits():
    spec = importlib.util.spec_from_file_location("_hq_circuits", ROOT / "hybridq_b200" / "circuits.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_hq_circuits"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    out_path, mode = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "plain")
    circuits = load_circuits()
    import hybridq.circuit.simulation.simulation as sim
    from hybridq.circuit import Circuit
    from hybridq.gate import MatrixGate

    # which file the bare name 'hybridq.so' resolved to
    mapped = sorted({line.split()[-1] for line in open("/proc/self/maps")
                     if "hybridq" in line.rsplit("/", 1)[-1] and ".so" in line})
    info = {"core": ";".join(mapped), "log2_pack_size": int(sim._log2_pack_size or 0)}
    assert sim._dot_core and sim._swap_core and sim._to_complex_core, "the reference did not find a native core"

    if mode == "dispatch":
        sys.path.insert(0, str(ROOT))
        import hybridq_b200
        real = sim._simulate_evolution
        calls = {"n": 0}

        def patched(circuit, initial_state, final_state, optimize, backend, complex_type, verbose, **kwargs):
            if optimize == "hybridq":
                calls["n"] += 1
                return hybridq_b200.simulate(circuit, initial_state=initial_state, optimize="evolution",
                                             complex_type=complex_type, simplify=False, compress=0,
                                             return_info=kwargs["return_info"],
                                             return_numpy_array=kwargs["return_numpy_array"])
            return real(circuit, initial_state, final_state, optimize, backend, complex_type, verbose, **kwargs)

        sim._simulate_evolution = patched

    n = 20
    gates = circuits.matching_circuit(n, depth=20, seed=20)
    circ = Circuit(MatrixGate(g.U, qubits=list(g.qubits)) for g in gates)
    out = {}
    for ctype in ("complex64", "complex128"):
        for tag, kw in (("c0", dict(simplify=False, compress=0)), ("c4", dict())):     # compress=4 is the default
            psi, sim_info = sim.simulate(circ, initial_state="0" * n, optimize="evolution", complex_type=ctype,
                                         return_info=True, max_largest_intermediate=2 ** 26, **kw)
            out[f"{ctype}_{tag}"] = np.asarray(psi).reshape(-1)
            out[f"{ctype}_{tag}_runtime"] = np.float64(sim_info["runtime (s)"])
    if mode == "dispatch":
        assert calls["n"] == 4, calls
    out["core"] = np.array(info["core"])
    out["log2_pack_size"] = np.int32(info["log2_pack_size"])
    np.savez(out_path, **out)
    print("REF_WORKER_OK", info, flush=True)


if __name__ == "__main__":
    main()
