"""Opcode histogram per kernel of libgplite_b200.so (cuobjdump -sass): the evidence of which hardware
paths the shipped binary uses -- DMMA.8x8x4 (FP64 tensor pipe; tcgen05 has no f64 kind on sm_100a),
LDGSTS (cp.async), MUFU.RSQ64H, and no HMMA / UTC*MMA / UTMA*.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "gplite_b200", "libgplite_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_x]+)*)", line)
    if m and kern:
        op = m.group(1)
        base = op.split(".")[0]
        key = op if base in ("DMMA", "MUFU", "HMMA", "LDGSTS", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM") else base
        hist[kern][key] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
total = collections.Counter()
print(f"# SASS opcode summary of {os.path.relpath(so, ROOT)} ({len(hist)} kernels), sm_100a")
for (k, h), name in zip(hist.items(), demangle):
    total.update(h)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)[:110]
    top = ", ".join(f"{o} {c}" for o, c in h.most_common(8))
    flags = [o for o in h if o.startswith(("DMMA", "HMMA", "UTC", "UTMA", "UBLKCP", "LDTM", "LDGSTS", "MUFU.RSQ64H"))]
    print(f"{name}\n    {sum(h.values())} instr: {top}\n    notable: {', '.join(f'{o} {h[o]}' for o in sorted(flags)) or '-'}")
print("\n# whole library")
for o in ("DMMA.8x8x4", "LDGSTS.E.BYPASS.128", "LDGSTS", "MUFU.RSQ64H", "HMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM"):
    print(f"{o}: {sum(c for k, c in total.items() if k.startswith(o))}")
print("UTC*MMA:", sum(c for k, c in total.items() if k.startswith("UTC") and "MMA" in k))
