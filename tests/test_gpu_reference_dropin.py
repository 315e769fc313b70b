"""The unmodified reference running over this repo's drop-in libraries, on the GPU box
(VERDICT r01 items "missing 3" / "weak 8").

oracle/_ref/pkg holds a git-ignored, unmodified copy of the reference's Python package and of its
tests/tests.py (made by `make -C oracle ref` from /root/reference; it travels to the GPU box like the compiled
reference core does).  Everything here runs the reference in a SUBPROCESS whose LD_LIBRARY_PATH puts
hybridq_b200/lib/dropin first, so `load_library('hybridq.so')` / `('hybridq_swap.so')`
(hybridq/utils/dot.py:35, transpose.py:34) bind the CUDA library.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from helpers import TOL, ROOT

pytestmark = pytest.mark.gpu

REFDIR = Path(ROOT) / "oracle" / "_ref"
DROPIN = Path(ROOT) / "hybridq_b200" / "lib" / "dropin"


def _env(libdir):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"{libdir}:" + env.get("LD_LIBRARY_PATH", "")
    env["PYTHONPATH"] = f"{REFDIR / 'pkg'}:{REFDIR / 'stubs'}"
    env.setdefault("OMP_NUM_THREADS", "8")
    return env


def _need_pkg():
    if not (REFDIR / "pkg" / "hybridq" / "__init__.py").exists() or not (REFDIR / "avx2" / "hybridq.so").exists():
        pytest.skip("oracle/_ref/pkg not present (run `make -C oracle ref` where /root/reference exists)")


def _run_worker(tmp_path, libdir, mode, name):
    out = tmp_path / f"{name}.npz"
    r = subprocess.run([sys.executable, "-W", "ignore", str(Path(ROOT) / "tests" / "ref_simulate_worker.py"), str(out), mode],
                       env=_env(libdir), cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "REF_WORKER_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
    return np.load(out)


def test_reference_simulate_over_dropin_config1(tmp_path):
    """Reference `simulate(optimize='evolution')`, n = 20 depth 20, complex64 / complex128, compress 0 and the
    default 4: once over the reference's own AVX2 core, once over the drop-in CUDA library (per-gate
    swap_* / apply_U_* / to_complex* calls, host pointers), once with the INTEGRATION.md section-2 dispatch
    (device-resident).  All three must agree to the north-star tolerance."""
    _need_pkg()
    ref = _run_worker(tmp_path, REFDIR / "avx2", "plain", "ref")
    drop = _run_worker(tmp_path, DROPIN, "plain", "dropin")
    disp = _run_worker(tmp_path, DROPIN, "dispatch", "dispatch")
    # which core each arm bound: this library reports a pack width of 2^1, the reference's AVX2 build 2^3
    assert int(drop["log2_pack_size"]) == 1 and int(disp["log2_pack_size"]) == 1, (str(drop["core"]), str(disp["core"]))
    assert int(ref["log2_pack_size"]) == 3 and "hybridq_b200" not in str(ref["core"]), str(ref["core"])
    for ctype in ("complex64", "complex128"):
        for tag in ("c0", "c4"):
            a = ref[f"{ctype}_{tag}"]
            assert abs(np.linalg.norm(a.astype(np.complex128)) - 1) < 1e-5
            assert np.abs(drop[f"{ctype}_{tag}"] - a).max() <= TOL[ctype], (ctype, tag, "drop-in")
            assert np.abs(disp[f"{ctype}_{tag}"] - a).max() <= TOL[ctype], (ctype, tag, "dispatch")


def test_reference_own_dot_and_transpose_tests_over_dropin(tmp_path):
    """The reference's own `test_utils__dot[...numpy...]` and `test_utils__transpose[...numpy...]`
    (/root/reference/tests/tests.py:304, :261; 500 parametrised cases, `raise_if_hcore_fails=True`) with the
    drop-in libraries bound.  Two pieces of scaffolding, both outside the reference files: an empty `cirq` import
    stub and a conftest that restores numpy.alltrue (removed in numpy 2)."""
    _need_pkg()
    r = subprocess.run([sys.executable, "-m", "pytest", str(REFDIR / "pkg" / "ref_tests" / "tests.py"), "-q", "-x",
                        "-p", "no:cacheprovider", "-k", "(test_utils__dot or test_utils__transpose) and numpy"],
                       env=_env(DROPIN), cwd=str(tmp_path), capture_output=True, text=True, timeout=1800)
    tail = r.stdout[-1500:] + r.stderr[-1500:]
    assert r.returncode == 0, tail
    assert "500 passed" in r.stdout, tail
