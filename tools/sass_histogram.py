"""Opcode histogram of every kernel in the built library (cuobjdump -sass), the evidence file
profiles/r02_sass_opcodes.txt: which kernels use the FP64 tensor cores (DMMA.8x8x4), the bulk-copy
engine (UBLKCP / UBLKPF) and mbarriers (SYNCS).  Runs without a GPU.
usage: python tools/sass_histogram.py [path-to-.so] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromevol_enhanced_b200 import _build  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else _build.LIB_PATH
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEY = ["DMMA.8x8x4", "HMMA", "UBLKCP", "UBLKPF", "SYNCS", "LDS", "STS", "LDG", "STG", "DADD", "DFMA", "DMUL",
       "BAR", "NANOSLEEP", "ATOM", "RED"]
kernels = collections.OrderedDict()
cur = None
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for k in KEY:
            if op == k or op.startswith(k + ".") or (k == "DMMA.8x8x4" and op.startswith("DMMA")):
                cur[k] += 1
print(f"# {os.path.basename(path)}: arch {sorted(arch)}; opcode counts per kernel (static SASS)")
print(f"# {'kernel':70s} {'instr':>6s} " + " ".join(f"{k.split('.')[0][:6]:>6s}" for k in KEY))
tot = collections.Counter()
for name, c in kernels.items():
    short = re.sub(r"\(int\)|\(bool\)", "", demangle(name)).replace("cevb::", "").replace("void ", "")
    short = re.sub(r"\(.*", "", short).replace(" ", "")
    print(f"{short[:72]:72s} {c['_total']:6d} " + " ".join(f"{c[k]:6d}" for k in KEY))
    tot.update(c)
print(f"{'TOTAL':72s} {tot['_total']:6d} " + " ".join(f"{tot[k]:6d}" for k in KEY))
