#!/usr/bin/env python
"""Instruction mix of the tile-kernel variants in the built library (cuobjdump -sass): how many HMMA / DMMA
(mma.sync tensor-core instructions), FFMA2 / FFMA / DFMA, shared-memory and cp.async (LDGSTS) instructions
each variant carries.  No GPU needed.  Usage: python tools/sass_mix.py > profiles/r01/sass_mix.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
lib = ROOT / "hybridq_b200" / "lib" / "libhybridq_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
fn = None
mix = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        mix[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        mix[fn][m.group(1)] += 1
keys = ["HMMA", "DMMA", "FFMA2", "FFMA", "DFMA", "FADD", "LDS", "STS", "LDGSTS", "LDG", "STG", "BAR", "LDL", "STL"]
print(f"# {lib.name}: static SASS instruction counts per kernel (cuobjdump -sass); HMMA = mma.sync TF32, DMMA = mma.sync FP64")
print(f"{'kernel':70s} " + " ".join(f"{k:>7s}" for k in keys) + "   total")
for fn, c in mix.items():
    if "hq_tile_kernel" not in fn and "hq_direct_kernel" not in fn and "gate_small_generic" not in fn:
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(.*", "", name).replace("void hq::", "")
    print(f"{name[:70]:70s} " + " ".join(f"{c.get(k, 0):7d}" for k in keys) + f" {sum(c.values()):7d}")
